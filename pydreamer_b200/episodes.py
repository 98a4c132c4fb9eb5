"""Episode files of the reference's replay buffer and the sequential batch iterator that feeds the training step
(the raw batches go through `pydreamer_b200.preprocess.GpuPreprocessor` on the device).

On-disk format (what `generator.py` writes and `data.py` reads):
  * one compressed `.npz` per episode (or chunk of episodes), written to memory first and then to the file
    (`tools.py:200-207`); keys `action, reward, terminal, reset, image | image_t, ...`, first axis = time;
  * RGB images are stored as `image_t` = THWC -> HWCT uint8 (`generator.py:246-249`, better compression) and
    transposed back on load (`data.py:236-239`);
  * file name `ep{from:06}_{to:06}[-{chunk}]-r{reward:.0f}-{steps:04}.npz` (`data.py:98-102`), parsed back into
    `(episode_from, episode_to, steps)` (`data.py:105-122`).

Batching (`data.py:127-305`, `DataSequential`): `batch_size` independent streams, each walking randomly chosen files
front to back in windows of `batch_length` steps (first window of the very first file starts at a random offset);
every file starts with `reset=True, reward=0`; optional random extra resets every ~`reset_interval` steps; the streams
are stacked to time-major `(T, B, ...)` arrays.  Random draws use the global `numpy.random` state in the same order as
the reference, so the same seed gives the same batches (tests/test_episodes.py pins that against the real reference).
"""
import io
import os
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------------------------- files
def episode_file_name(episode_from: int, episode_to: int, reward: float, steps: int, chunk_seq: Optional[int] = None) -> str:
    """data.py:98-102"""
    if chunk_seq is None:
        return f"ep{episode_from:06}_{episode_to:06}-r{reward:.0f}-{steps:04}.npz"
    return f"ep{episode_from:06}_{episode_to:06}-{chunk_seq}-r{reward:.0f}-{steps:04}.npz"


def parse_episode_name(fname: str) -> Tuple[int, int, int]:
    """(episode_from, episode_to, steps) from a file name; unknown pieces parse as 0 (data.py:105-122)."""
    stem = fname.split("/")[-1].split(".")[0]
    steps = stem.split("-")[-1]
    steps = int(steps) if steps.isnumeric() else 0
    if not stem.startswith("ep"):
        return 0, 0, steps                                # '{timestamp}-{steps}'
    span = stem.split("ep")[1].split("-")[0]
    lo, hi = span.split("_")[0], span.split("_")[-1]
    return (int(lo) if lo.isnumeric() else 0), (int(hi) if hi.isnumeric() else 0), steps


def save_episode(data: Dict[str, np.ndarray], directory, episode_from: int, episode_to: int,
                 chunk_seq: Optional[int] = None) -> Path:
    """Write one episode (chunk) the way generator.py:244-251 + data.py:62-68 do; returns the path."""
    data = dict(data)
    if "image" in data and data["image"].ndim == 4:
        data["image_t"] = data.pop("image").transpose(1, 2, 3, 0)          # THWC -> HWCT
    n_episodes = int(data["reset"].sum())
    n_steps = len(data["reset"]) - n_episodes
    path = Path(directory) / episode_file_name(episode_from, episode_to, float(data["reward"].sum()), n_steps, chunk_seq)
    path.parent.mkdir(parents=True, exist_ok=True)
    with io.BytesIO() as buf:                                               # tools.py:200-207
        np.savez_compressed(buf, **data)
        buf.seek(0)
        with path.open("wb") as f:
            f.write(buf.read())
    return path


def load_episode(path) -> Dict[str, np.ndarray]:
    """tools.py:210-219 + the fix-ups of data.py:236-256: image_t -> image (THWC), `action_next`, the file starts
    with a reset and carries no reward on its first step."""
    with Path(path).open("rb") as f:
        z = np.load(f)
        data = {k: z[k] for k in z}
    if "image" not in data and "image_t" in data:
        data["image"] = data.pop("image_t").transpose(3, 0, 1, 2)          # HWCT -> THWC
    data["action_next"] = np.concatenate([data["action"][1:], np.zeros_like(data["action"][:1])])
    n = data["reward"].shape[0]
    if "reset" not in data:
        data["reset"] = np.zeros(n, bool)
    data["reset"][0] = True
    data["reward"][0] = 0.0
    return data


@dataclass
class EpisodeFile:
    path: str
    episode_from: int
    episode_to: int
    steps: int

    def load_data(self) -> Dict[str, np.ndarray]:
        with Path(self.path).open("rb") as f:
            z = np.load(f)
            return {k: z[k] for k in z}


class EpisodeDirectory:
    """A directory of episode files (the local-filesystem case of data.py:54-95)."""

    def __init__(self, directory):
        self.directory = Path(directory)

    def save_data(self, data, episode_from, episode_to, chunk_seq=None):
        return save_episode(data, self.directory, episode_from, episode_to, chunk_seq)

    def list_files(self) -> List[EpisodeFile]:
        files = []
        for name in sorted(os.listdir(self.directory)):
            if name.endswith(".npz"):
                lo, hi, steps = parse_episode_name(name)
                files.append(EpisodeFile(str(self.directory / name), lo, hi, steps))
        return files

    def count_steps(self):
        files = self.list_files()
        return len(files), sum(f.steps for f in files), (max(f.episode_to for f in files) + 1) if files else 0


# --------------------------------------------------------------------------------------------------------- batches
def _len(batch) -> int:
    return batch["reward"].shape[0]


class SequentialBatches:
    """data.py:127-305 (`DataSequential`) over an `EpisodeDirectory`-like object: iterating yields time-major dicts
    of numpy arrays `(batch_length, batch_size, ...)` forever."""

    def __init__(self, repository, batch_length, batch_size, skip_first=True, reload_interval=0, buffer_size=0,
                 reset_interval=0, allow_mid_reset=False, check_nonempty=True):
        self.repository = repository
        self.batch_length, self.batch_size = batch_length, batch_size
        self.skip_first, self.reload_interval, self.buffer_size = skip_first, reload_interval, buffer_size
        self.reset_interval, self.allow_mid_reset = reset_interval, allow_mid_reset
        self.reload_files()
        if check_nonempty:
            assert len(self.files) > 0, "No data found"

    def reload_files(self):
        files_all = self.repository.list_files()
        files_all.sort(key=lambda e: -e.episode_to)                        # newest first (data.py:160)
        files, total = [], 0
        for f in files_all:
            total += f.steps
            if total < self.buffer_size or not self.buffer_size:
                files.append(f)
        self.files, self.last_reload, self.stats_steps = files, time.time(), total

    def should_reload_files(self):
        return self.reload_interval and (time.time() - self.last_reload > self.reload_interval)

    def __iter__(self) -> Iterator[Dict[str, np.ndarray]]:
        streams = [self._stream() for _ in range(self.batch_size)]
        for windows in zip(*streams):                                       # one window per stream, in stream order
            keys = set(windows[0])
            for w in windows[1:]:
                keys &= set(w)
            yield {k: np.stack([w[k] for w in windows]).swapaxes(0, 1) for k in keys}

    def _shuffled_files(self):
        while True:
            if self.should_reload_files():
                self.reload_files()
            yield self.files[np.random.choice(len(self.files))]

    def _stream(self):
        skip_random = self.skip_first
        carry = None                                                        # partial window waiting for the next file
        for file in self._shuffled_files():
            first_len = self.batch_length - _len(carry) if carry else None
            it = self._file_windows(file, skip_random, first_len)
            if carry is not None:
                for window, partial in it:
                    assert not partial, "First batch must be full. Is episode_length < batch_size?"
                    keys = set(carry) & set(window)
                    window = {k: np.concatenate([carry[k], window[k]]) for k in keys}
                    assert _len(window) == self.batch_length
                    carry = None
                    yield window
                    break
            for window, partial in it:
                if partial:
                    carry = window if self.allow_mid_reset else None
                    break
                yield window
            skip_random = False

    def _file_windows(self, file, skip_random, first_shorter_length):
        try:
            data = file.load_data()
        except Exception as e:                                              # unreadable file: skip it (data.py:228-234)
            print("Error reading file - skipping")
            print(e)
            return
        if "image" not in data and "image_t" in data:
            data["image"] = data.pop("image_t").transpose(3, 0, 1, 2)
        data["action_next"] = np.concatenate([data["action"][1:], np.zeros_like(data["action"][:1])])
        n = _len(data)
        if n < self.batch_length:
            print(f"Skipping too short file: {file.path}, len={n}")
            return
        if "reset" not in data:
            data["reset"] = np.zeros(n, bool)
        data["reset"][0] = True
        data["reward"][0] = 0.0
        i = 0 if not skip_random else np.random.randint(n - self.batch_length + 1)
        length = first_shorter_length or self.batch_length
        if self.reset_interval:
            random_resets = self._randomize_resets(data["reset"], self.reset_interval, self.batch_length)
        else:
            random_resets = np.zeros_like(data["reset"])
        while i < n:
            window = {k: v[i:i + length] for k, v in data.items()}
            if np.any(random_resets[i:i + length]):
                assert not np.any(window["reset"]), "randomize_resets should not coincide with actual resets"
                window["reset"][0] = True                                   # views: this marks `data` too, as upstream
            partial = _len(window) < length
            i += length
            length = self.batch_length
            yield window, partial

    @staticmethod
    def _randomize_resets(resets, reset_interval, batch_length):
        """data.py:284-304: cut every episode into a random number of intervals, at least batch_length apart."""
        assert resets[0]
        bounds = np.where(resets)[0].tolist() + [len(resets)]
        out = np.zeros_like(resets)
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            steps = hi - lo
            n_int = np.random.randint(1, steps // reset_interval + 2)
            cuts = np.sort(np.random.choice(steps - batch_length * n_int, n_int - 1))
            cuts = lo + cuts + np.arange(1, n_int) * batch_length
            out[cuts] = True
            assert (resets | out)[lo:hi].sum() == n_int
        return out
