"""BASELINE.json configs[0]: `--configs defaults minigrid`, batch 10 x seq 10, one `Dreamer.training_step` on a synthetic
replay batch on CPU — the reference's own CPU-runnable plumbing case.  This configuration (dense image encoder / decoder,
categorical 7x7 observations, reward input, map probe) is OUTSIDE the accelerated hot path (§8f):
the test pins what a user of that config gets —
  * the drop-in module refuses it loudly at construction (no silent partial support),
  * the unmodified reference runs it end to end with the batch contract of SURVEY.md §8(d) (needs the reference importable:
    /root/reference in the authoring container or baseline/_ref), with the parameter count and key sets recorded in the survey.
"""
import os
import sys
from argparse import Namespace

import pytest
import torch

from pydreamer_b200.config import DEFAULTS
from pydreamer_b200.dreamer import Dreamer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# config/defaults.yaml:122-141 (`minigrid` section), hot-path keys
MINIGRID = dict(image_key="image", image_size=7, image_channels=4, image_categorical=True, map_key="map", map_size=11,
                map_channels=4, map_categorical=True, action_dim=7, reward_input=True, image_encoder="dense",
                image_encoder_layers=3, image_decoder="dense", image_decoder_layers=2, probe_model="map", imag_horizon=1)


def minigrid_conf(**over):
    d = dict(DEFAULTS)
    d.update(MINIGRID)
    d.update(batch_size=10, batch_length=10, device="cpu")
    d.update(over)
    return Namespace(**d)


def minigrid_batch(conf, seed=0):
    """image one-hot (T,B,4,7,7), map one-hot (T,B,4,11,11), map_coord (T,B,4), one-hot actions."""
    g = torch.Generator().manual_seed(seed)
    T, B, A = conf.batch_length, conf.batch_size, conf.action_dim
    oh = lambda n, c, s: torch.nn.functional.one_hot(torch.randint(0, c, (T, B, s, s), generator=g), c).permute(0, 1, 4, 2, 3).float()
    reset = torch.rand(T, B, generator=g) < 0.005
    reset[0] = True
    return dict(image=oh(T, conf.image_channels, conf.image_size), map=oh(T, conf.map_channels, conf.map_size),
                map_coord=torch.rand(T, B, 4, generator=g), map_seen_mask=torch.ones(T, B, conf.map_size, conf.map_size),
                action=torch.nn.functional.one_hot(torch.randint(0, A, (T, B), generator=g), A).float(),
                reward=torch.tanh(torch.randn(T, B, generator=g)), terminal=(torch.rand(T, B, generator=g) < 0.01).float(),
                reset=reset, vecobs=torch.zeros(T, B, 64))


def test_dropin_module_refuses_the_minigrid_config():
    with pytest.raises(NotImplementedError):
        Dreamer(minigrid_conf())


def test_reference_runs_the_minigrid_plumbing_step():
    RefDreamer = None
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "pydreamer")):
            sys.path.insert(0, cand)
            try:
                from pydreamer.models import Dreamer as RefDreamer
                break
            except Exception:
                continue
    if RefDreamer is None:
        pytest.skip("reference not importable here")
    torch.distributions.Distribution.set_default_validate_args(False)              # train.py:30
    conf = minigrid_conf()
    torch.manual_seed(0)
    model = RefDreamer(conf)
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 41_857_250
    obs = minigrid_batch(conf)
    state = model.init_state(conf.batch_size * conf.iwae_samples)
    losses, out_state, metrics, tensors, dream = model.training_step(obs, state)
    assert len(losses) == 4 and all(torch.isfinite(l).all() for l in losses)
    opts = model.init_optimizers(conf.adam_lr, conf.adam_lr_actor, conf.adam_lr_critic, conf.adam_eps)
    for o in opts:
        o.zero_grad()
    for l in losses:
        l.backward()
    norms = model.grad_clip(conf.grad_clip, conf.grad_clip_ac)
    for o in opts:
        o.step()
    assert set(norms) == {"grad_norm", "grad_norm_probe", "grad_norm_actor", "grad_norm_critic"}
    assert {"loss_model", "loss_kl", "entropy_prior", "entropy_post", "loss_actor", "loss_critic", "policy_value"} <= set(metrics)
    assert out_state[0].shape == (conf.batch_size, conf.deter_dim) and not out_state[0].requires_grad
    assert float(norms["grad_norm_probe"]) > 0                                     # the map probe head trains
