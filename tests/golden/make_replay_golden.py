"""Generates tests/golden/replay/* from the UNMODIFIED reference data pipeline (pydreamer/data.py, pydreamer/tools.py).

Run in the authoring container only (needs /root/reference):   python tests/golden/make_replay_golden.py
`pydreamer.data` imports mlflow at module level (absent here and irrelevant to the file format), so empty stand-in modules
are registered for the import; every function exercised below (tools.save_npz / load_npz, DataSequential and its helpers)
is the reference's own code.  Writes:
  replay/ep*.npz        four small synthetic episode files written with the reference's save_npz + generator.py:246-249
  replay_expected.npz   the first batches the reference's DataSequential yields for three seeded configurations
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "replay")
sys.path.insert(0, "/root/reference")
for name in ("mlflow", "mlflow.store", "mlflow.store.artifact", "mlflow.store.artifact.artifact_repo",
             "mlflow.store.artifact.artifact_repository_registry", "mlflow.tracking", "mlflow.tracking.client"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["mlflow.store.artifact.artifact_repo"].ArtifactRepository = object
sys.modules["mlflow.store.artifact.artifact_repository_registry"].get_artifact_repository = lambda uri: None
sys.modules["mlflow.tracking"].MlflowClient = object
sys.modules["mlflow.tracking.client"].MlflowClient = object

from pydreamer import data as ref_data      # noqa: E402
from pydreamer import tools as ref_tools    # noqa: E402

CONFIGS = dict(plain=dict(batch_length=6, batch_size=3, skip_first=True, reset_interval=0, allow_mid_reset=False, seed=11, n=5),
               resets=dict(batch_length=5, batch_size=2, skip_first=True, reset_interval=12, allow_mid_reset=False, seed=12, n=6),
               midreset=dict(batch_length=7, batch_size=4, skip_first=False, reset_interval=0, allow_mid_reset=True, seed=13, n=6))


def synth_episode(rng, steps, A=5):
    return dict(image=rng.integers(0, 256, (steps, 8, 8, 3), dtype=np.uint8),
                action=np.eye(A, dtype=np.float32)[rng.integers(0, A, steps)],
                reward=rng.normal(size=steps).astype(np.float32),
                terminal=np.zeros(steps, bool), reset=np.zeros(steps, bool))


class DirRepo(ref_data.MlflowEpisodeRepository):
    """list_files over a plain directory; names are built / parsed by the reference's own methods."""

    def __init__(self, directory):
        self.directory = directory
        self.artifact_uris = [directory]

    def list_files(self):
        files = []
        for name in sorted(os.listdir(self.directory)):
            if name.endswith(".npz"):
                lo, hi, steps = self.parse_episode_name(name)
                files.append(LocalFile(os.path.join(self.directory, name), lo, hi, steps, None))
        return files


class LocalFile(ref_data.FileInfo):
    def load_data(self):
        return ref_tools.load_npz(self.path)


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.remove(os.path.join(OUT, f))
    rng = np.random.default_rng(2024)
    repo = DirRepo(OUT)
    for ep, steps in enumerate((41, 33, 57, 29)):
        d = synth_episode(rng, steps)
        d["reset"][0] = True
        d["terminal"][-1] = True
        d["image_t"] = d.pop("image").transpose(1, 2, 3, 0)                       # generator.py:246-249
        n_steps = len(d["reset"]) - d["reset"].sum()
        name = repo.build_episode_name(ep, ep, d["reward"].sum(), n_steps)         # data.py:62-66
        ref_tools.save_npz(d, os.path.join(OUT, name))
    expected = {}
    for cname, c in CONFIGS.items():
        np.random.seed(c["seed"])
        ds = ref_data.DataSequential(repo, c["batch_length"], c["batch_size"], skip_first=c["skip_first"],
                                     reset_interval=c["reset_interval"], allow_mid_reset=c["allow_mid_reset"])
        it = iter(ds)
        for i in range(c["n"]):
            batch = next(it)
            for k, v in batch.items():
                expected[f"{cname}/{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "replay_expected.npz"), **expected)
    print("files:", sorted(os.listdir(OUT)), "arrays:", len(expected))


if __name__ == "__main__":
    main()
