"""The oracle restatement against the committed golden vectors (outputs of the unmodified reference,
tests/golden/make_golden.py).  fp32, CPU; tolerance 2e-4 relative (different op order than the reference's
nn.Modules; the generator script measured <= 1e-5 on gradients)."""
import pytest
import torch

from oracle import dreamer_oracle as O
from pydreamer_b200.dreamer import Dreamer
from tests.util import CASES, build_case, seeded_weights


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_golden(case):
    fx, conf, obs, state, noise = build_case(case)
    template = Dreamer(conf).state_dict()            # same keys/shapes as the reference's state_dict
    sd = {k: v.clone().requires_grad_(not k.startswith("ac.critic_target")) for k, v in seeded_weights(template, fx).items()}
    res = O.training_step(sd, conf, obs, state, noise)
    for l in res["losses"]:
        l.backward()
    for got, want in zip(res["losses"], fx["losses"]):
        assert abs(float(got.detach().reshape(-1)[0]) - want) <= 2e-5 * max(1.0, abs(want))
    for k, want in fx["metrics"].items():
        assert abs(float(res["metrics"][k]) - want) <= 2e-4 * max(1.0, abs(want)), k
    assert set(fx["grad_norms"]) == {k for k, v in sd.items() if v.grad is not None}
    for k, want in fx["grad_norms"].items():
        got = float(sd[k].grad.double().norm())
        assert abs(got - want) <= 2e-4 * max(want, 1e-6) + 1e-9, (k, got, want)
    for k, want in fx["tensor_abs_sums"].items():
        got = float(res["tensors"][k].double().abs().sum())
        assert abs(got - want) <= 2e-4 * max(want, 1e-6), (k, got, want)
    assert int(res["inter"]["post_idx"].sum()) == fx["post_sample_index_sum"]
    assert res["inter"]["post_idx"][0].reshape(-1).tolist() == fx["post_sample_indices_t0"]   # bit-exact indices
    assert abs(float(res["inter"]["dream_actions"].double().sum()) - fx["dream_action_sum"]) < 1e-4
