"""SURVEY.md §8(f) N1 on the GPU: `Learner.step` through the native kernels (CUDA-graph steady state included), TBTT state
carry, logging step, checkpoint round trip in the reference's format (model + torch-layout optimizer state)."""
import pytest
import torch

from pydreamer_b200.config import make_conf
from pydreamer_b200.learner import Learner
from pydreamer_b200.replay import synthetic_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_learner_steps_on_gpu_and_checkpoint_roundtrip(tmp_path):
    conf = make_conf("tiny", device=DEV)
    lr = Learner(conf, DEV)
    batches = [synthetic_batch(conf, seed=s, first=(s == 1)) for s in (1, 2, 3, 4, 5)]        # host batches (train.py:160-161)
    before = lr.model.wm.core.cell.z_mlp.weight.detach().clone()
    seen = []
    for i, b in enumerate(batches):                      # calls 1-2 eager, call 3 captures the CUDA graph, 4-5 replay it
        met, tensors, dream = lr.step(b, do_image_pred=False)
        assert {"loss_model", "loss_kl", "grad_norm", "grad_norm_actor", "grad_norm_critic", "grad_norm_probe"} <= set(met)
        seen.append({k: float(v) for k, v in met.items()})
        assert all(v == v for v in seen[-1].values()), seen[-1]          # finite
    assert lr.steps == 5 and 0 in lr.states and lr.states[0][0].shape == (conf.batch_size, conf.deter_dim)
    assert any(g.get("graph") is not None for g in lr.model._graphs.values())                 # the steady state replays a graph
    assert not torch.equal(before, lr.model.wm.core.cell.z_mlp.weight.detach())
    assert seen[0]["loss_model"] != seen[-1]["loss_model"]
    met, tensors, dream = lr.step(batches[0], do_image_pred=True, do_dream_tensors=True)      # a logging step (eager)
    assert "logprob_image" in met and "image_pred" in tensors and "value" in dream
    path = str(tmp_path / "latest.pt")
    lr.save_checkpoint(path)
    ck = torch.load(path)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_0_state_dict", "optimizer_1_state_dict",
                       "optimizer_2_state_dict", "optimizer_3_state_dict"}                      # tools.py:164-174
    assert float(ck["optimizer_0_state_dict"]["state"][0]["step"]) == 6.0
    lr2 = Learner(conf, DEV)
    assert lr2.load_checkpoint(path) == 6
    for (k, a), (_, b_) in zip(lr.model.state_dict().items(), lr2.model.state_dict().items()):
        assert torch.equal(a, b_), k
    for o1, o2 in zip(lr.optimizers, lr2.optimizers):
        assert torch.equal(o1.exp_avg, o2.exp_avg) and torch.equal(o1.exp_avg_sq, o2.exp_avg_sq) and int(o2.step_t) == 6
    m2, _, _ = lr2.step(batches[1])                         # the restored learner trains on
    assert float(m2["grad_norm"]) > 0 and float(m2["loss_model"]) == float(m2["loss_model"])
