"""Replay episode files + sequential batching (pydreamer_b200/episodes.py) against fixtures produced by the UNMODIFIED
reference pipeline (tests/golden/make_replay_golden.py: pydreamer/data.py DataSequential over files written with
pydreamer/tools.py save_npz).  Bit-exact: same files, same numpy seed => same batches."""
import os

import numpy as np

from pydreamer_b200 import episodes as E

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay")
CONFIGS = dict(plain=dict(batch_length=6, batch_size=3, skip_first=True, reset_interval=0, allow_mid_reset=False, seed=11, n=5),
               resets=dict(batch_length=5, batch_size=2, skip_first=True, reset_interval=12, allow_mid_reset=False, seed=12, n=6),
               midreset=dict(batch_length=7, batch_size=4, skip_first=False, reset_interval=0, allow_mid_reset=True, seed=13, n=6))


def test_file_names_round_trip_like_the_reference():
    files = E.EpisodeDirectory(GOLD).list_files()
    assert [os.path.basename(f.path) for f in files] == ["ep000000_000000-r-12-0040.npz", "ep000001_000001-r-3-0032.npz",
                                                         "ep000002_000002-r4-0056.npz", "ep000003_000003-r2-0028.npz"]
    assert [(f.episode_from, f.episode_to, f.steps) for f in files] == [(0, 0, 40), (1, 1, 32), (2, 2, 56), (3, 3, 28)]
    assert E.episode_file_name(12, 14, -3.4, 1999, chunk_seq=1) == "ep000012_000014-1-r-3-1999.npz"
    assert E.parse_episode_name("some/dir/ep000012_000014-1-r-3-1999.npz") == (12, 14, 1999)
    assert E.parse_episode_name("20210101T000000-0500.npz") == (0, 0, 500)
    assert E.EpisodeDirectory(GOLD).count_steps() == (4, 156, 4)


def test_reference_written_file_loads_and_our_writer_is_readable_the_same_way(tmp_path):
    src = E.EpisodeDirectory(GOLD).list_files()[2]
    ep = E.load_episode(src.path)
    assert ep["image"].shape == (57, 8, 8, 3) and ep["image"].dtype == np.uint8            # HWCT on disk -> THWC
    assert ep["reset"][0] and ep["reward"][0] == 0.0 and ep["action_next"].shape == ep["action"].shape
    assert np.array_equal(ep["action_next"][:-1], ep["action"][1:]) and not ep["action_next"][-1].any()
    raw = src.load_data()
    path = E.save_episode(dict(image=raw["image_t"].transpose(3, 0, 1, 2), action=raw["action"], reward=raw["reward"],
                               terminal=raw["terminal"], reset=raw["reset"]), tmp_path, 2, 2)
    assert os.path.basename(path) == os.path.basename(src.path)                            # same name schema
    again = E.EpisodeFile(str(path), 2, 2, 56).load_data()
    assert set(again) == set(raw) and all(np.array_equal(again[k], raw[k]) for k in raw)   # same stored arrays (image_t)


def test_sequential_batches_equal_the_reference_iterator_bit_for_bit():
    want = np.load(os.path.join(os.path.dirname(GOLD), "replay_expected.npz"))
    for name, c in CONFIGS.items():
        np.random.seed(c["seed"])
        ds = E.SequentialBatches(E.EpisodeDirectory(GOLD), c["batch_length"], c["batch_size"], skip_first=c["skip_first"],
                                 reset_interval=c["reset_interval"], allow_mid_reset=c["allow_mid_reset"])
        it = iter(ds)
        for i in range(c["n"]):
            batch = next(it)
            keys = {k.split("/")[2] for k in want.files if k.startswith(f"{name}/{i}/")}
            assert set(batch) == keys
            for k in keys:
                w = want[f"{name}/{i}/{k}"]
                assert batch[k].shape == w.shape and batch[k].dtype == w.dtype, (name, i, k)
                assert np.array_equal(batch[k], w), (name, i, k)
            assert batch["reward"].shape == (c["batch_length"], c["batch_size"])           # time-major (T, B)


def test_batches_feed_the_preprocessor_and_the_module_contract():
    """episode files -> SequentialBatches -> preprocessing -> the obs dict Dreamer.training_step takes."""
    from oracle import preprocess_oracle as P
    from pydreamer_b200.config import make_conf

    conf = make_conf("tiny")
    np.random.seed(5)
    batch = next(iter(E.SequentialBatches(E.EpisodeDirectory(GOLD), conf.batch_length, conf.batch_size)))
    obs = P.apply(batch, conf.action_dim)
    T, B = conf.batch_length, conf.batch_size
    assert obs["image"].shape == (T, B, 3, 8, 8) and obs["image"].dtype == np.float32
    assert obs["image"].min() >= -0.5 and obs["image"].max() <= 0.5
    assert obs["action"].shape == (T, B, conf.action_dim) and obs["reset"].dtype == bool
    assert obs["reward"].shape == obs["terminal"].shape == (T, B)
