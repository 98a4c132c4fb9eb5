"""Data-parallel host logic on CPU: world_size 2 over gloo, each rank runs the module (reference op table) on its
half of the global batch; after the single flat-bucket all-reduce the gradients, clipped norms and the updated
parameters must equal a single process running the concatenated global batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dreamer_oracle as O
from oracle.ref_ops import RefOps
from oracle.weights import seeded_state_dict
from pydreamer_b200 import ops as pd_ops
from pydreamer_b200.config import make_conf
from pydreamer_b200.dreamer import Dreamer
from pydreamer_b200.parallel import GradAllReduce
from pydreamer_b200.replay import synthetic_batch

WORLD, BG = 2, 4


def _inputs():
    conf = make_conf("tiny", device="cpu", batch_size=BG)
    obs = synthetic_batch(conf, seed=5)
    torch.manual_seed(11)
    noise = O.draw_noise(conf, conf.batch_length, BG)
    return conf, obs, noise


def _shard(conf, obs, noise, r):
    T, Bl = conf.batch_length, BG // WORLD
    sl = slice(r * Bl, (r + 1) * Bl)
    o = {k: v[:, sl].contiguous() for k, v in obs.items()}
    H = conf.imag_horizon
    n = dict(post=noise["post"][:, sl].contiguous(),
             actor=noise["actor"].view(H, T, BG, -1)[:, :, sl].reshape(H, T * Bl, -1).contiguous(),
             prior=noise["prior"].view(H, T, BG, -1)[:, :, sl].reshape(H, T * Bl, -1).contiguous())
    return o, n


def _run(model, conf, obs, noise, B):
    opts = model.init_optimizers(conf.adam_lr, conf.adam_lr_actor, conf.adam_lr_critic, conf.adam_eps)
    losses, *_ = model.training_step(obs, model.init_state(B), noise=noise)
    for l in losses:
        l.backward()
    norms = model.grad_clip(conf.grad_clip, conf.grad_clip_ac)
    for o in opts:
        o.step()
    return {k: float(v) for k, v in norms.items()}


def _worker(rank, port, out):
    os.environ["PD_B200_TESTING"] = "1"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=WORLD)
    pd_ops.set_ops_for_testing(RefOps("cpu"))
    conf, obs, noise = _inputs()
    lconf = make_conf("tiny", device="cpu", batch_size=BG // WORLD)
    model = Dreamer(lconf)
    model.fp16_forward = False
    if rank == 0:
        model.load_state_dict(seeded_state_dict(model.state_dict(), 3))
    model._dp = GradAllReduce(WORLD)
    model._dp.broadcast_params(model)
    o, n = _shard(conf, obs, noise, rank)
    norms = _run(model, lconf, o, n, BG // WORLD)
    if rank == 0:
        torch.save(dict(norms=norms, params={k: v.clone() for k, v in model.state_dict().items()}), out)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process_global_batch(tmp_path):
    out = str(tmp_path / "dp.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(port, out), nprocs=WORLD, join=True)
    got = torch.load(out)
    pd_ops.set_ops_for_testing(RefOps("cpu"))
    try:
        conf, obs, noise = _inputs()
        model = Dreamer(conf)
        model.fp16_forward = False
        model.load_state_dict(seeded_state_dict(model.state_dict(), 3))
        norms = _run(model, conf, obs, noise, BG)
    finally:
        pd_ops.set_ops_for_testing(None)
    for k, v in norms.items():
        assert abs(got["norms"][k] - v) <= 1e-4 * max(abs(v), 1e-6), (k, got["norms"][k], v)
    for k, v in model.state_dict().items():
        assert torch.allclose(got["params"][k], v, rtol=1e-4, atol=1e-6), k
