"""Shared helpers: rebuild the seeded inputs of a committed golden case (no reference checkout needed)."""
import json
import os

import torch

from oracle import dreamer_oracle as O
from oracle.weights import seeded_state_dict
from pydreamer_b200.config import make_conf
from pydreamer_b200.replay import synthetic_batch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("tiny_onehot", "tiny_iwae3", "tiny_dmc", "tiny_klbal05")


def load_fixture(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        return json.load(f)


def build_case(name, device="cpu"):
    """-> (fixture, conf, obs, in_state, noise, state_dict_template_fn)"""
    fx = load_fixture(name)
    conf = make_conf(fx["preset"], device=str(device), **fx["overrides"])
    T, B, I = conf.batch_length, conf.batch_size, conf.iwae_samples
    obs = synthetic_batch(conf, seed=fx["seeds"]["data"])
    g = torch.Generator().manual_seed(fx["seeds"]["state"])
    D, Z = conf.deter_dim, conf.stoch_dim * conf.stoch_discrete
    state = (torch.tanh(torch.randn((B * I, D), generator=g)), torch.zeros(B * I, Z))
    torch.manual_seed(fx["seeds"]["noise"])
    noise = O.draw_noise(conf, T, B)
    mv = lambda d: {k: v.to(device) for k, v in d.items()}
    return fx, conf, mv(obs), tuple(s.to(device) for s in state), mv(noise)


def seeded_weights(model_state_dict, fx):
    return seeded_state_dict(model_state_dict, fx["seeds"]["weights"])


def rel_err(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
