"""TEST INFRASTRUCTURE — numpy restatement of pydreamer/preprocessing.py for the hot-path keys
(`to_image` :21-29, `to_onehot` use at :135-138, reward/terminal :148-150, `clip_rewards_np` functions.py:153-160).
Checked against the real Preprocessor in tests/test_preprocess_cpu.py when the reference is reachable."""
import numpy as np


def apply(batch, action_dim, clip_rewards="tanh"):
    out = {}
    x = batch["image"]
    if x.dtype == np.uint8:
        x = x.astype(np.float32) / 255.0 - 0.5
    else:
        x = x.astype(np.float32)
    out["image"] = x.transpose(0, 1, 4, 2, 3)
    a = batch["action"]
    if a.ndim == 2:
        a = np.eye(action_dim, dtype=np.float32)[a]
    out["action"] = a.astype(np.float32)
    T, B = batch["reward"].shape[:2]
    out["terminal"] = batch.get("terminal", np.zeros((T, B))).astype(np.float32)
    r = batch.get("reward", np.zeros((T, B))).astype(np.float32)
    out["reward"] = np.tanh(r) if clip_rewards == "tanh" else r
    out["reset"] = batch["reset"].astype(bool)
    return out
