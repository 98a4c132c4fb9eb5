"""Learner step + checkpoint format on CPU with the reference op table; when a reference checkout /
install is reachable the checkpoint is also loaded, strict, into the UNMODIFIED reference Dreamer."""
import os
import sys

import pytest
import torch

from oracle.ref_ops import RefOps
from pydreamer_b200 import ops as pd_ops
from pydreamer_b200.config import make_conf
from pydreamer_b200.learner import Learner
from pydreamer_b200.replay import synthetic_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def ref_ops():
    pd_ops.set_ops_for_testing(RefOps("cpu"))
    yield
    pd_ops.set_ops_for_testing(None)


def test_learner_steps_carry_state_and_checkpoint_roundtrip(ref_ops, tmp_path):
    conf = make_conf("tiny", device="cpu")
    lr = Learner(conf, "cpu")
    m0 = {k: v.clone() for k, v in lr.model.state_dict().items()}
    b1, b2 = synthetic_batch(conf, seed=1), synthetic_batch(conf, seed=2, first=False)
    met1, tensors, _ = lr.step(b1)
    assert {"loss_model", "loss_kl", "grad_norm", "grad_norm_actor", "grad_norm_critic", "grad_norm_probe"} <= set(met1)
    assert 0 in lr.states and lr.states[0][0].shape == (conf.batch_size, conf.deter_dim)
    met2, _, _ = lr.step(b2, do_image_pred=True)
    assert "logprob_image" in met2 and lr.steps == 2
    changed = sum(not torch.equal(m0[k], v) for k, v in lr.model.state_dict().items())
    assert changed > 100                                      # parameters moved
    path = str(tmp_path / "latest.pt")
    lr.save_checkpoint(path)
    ck = torch.load(path)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_0_state_dict", "optimizer_1_state_dict",
                       "optimizer_2_state_dict", "optimizer_3_state_dict"}                       # tools.py:164-174
    lr2 = Learner(conf, "cpu")
    assert lr2.load_checkpoint(path) == 2
    for (k, a), (_, b_) in zip(lr.model.state_dict().items(), lr2.model.state_dict().items()):
        assert torch.equal(a, b_), k
    # the reference module, if reachable, must load the checkpoint strictly (generator.py:109)
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "pydreamer")):
            sys.path.insert(0, cand)
            try:
                from pydreamer.models import Dreamer as RefDreamer
            except Exception:
                continue
            ref = RefDreamer(conf)
            ref.load_state_dict(ck["model_state_dict"], strict=True)
            out = ref.training_step(b1, ref.init_state(conf.batch_size))
            assert torch.isfinite(out[0][0])
            break


def test_learner_trains_from_raw_replay_batches(ref_ops):
    """Raw replay batches (uint8 HWC images, integer actions: what the reference's DataSequential yields) -> device-side
    preprocessing -> gradient steps, state carried across consecutive windows."""
    import numpy as np

    conf = make_conf("tiny", device="cpu", reset_interval=0)
    rng = np.random.default_rng(0)
    T, B = conf.batch_length, conf.batch_size

    def batches():
        first = True
        while True:
            reset = np.zeros((T, B), bool)
            reset[0] = first
            first = False
            yield dict(image=rng.integers(0, 256, (T, B, 64, 64, 3), dtype=np.uint8),
                       action=rng.integers(0, conf.action_dim, (T, B)), reward=rng.normal(size=(T, B)).astype(np.float32),
                       terminal=np.zeros((T, B), bool), reset=reset)

    lr = Learner(conf, "cpu")
    metrics = lr.train_on_batches(batches(), 3)
    assert lr.steps == 3 and torch.isfinite(metrics["loss_model"]) and float(metrics["grad_norm"]) > 0
    assert lr.states[0][0].shape == (conf.batch_size, conf.deter_dim)          # state carried across the windows
