"""Runs a few eager gradient steps of the benchmark workload and brackets the LAST one with cudaProfilerStart/Stop,
so `ncu --profile-from-start off` sees exactly one steady-state step (launch list / --set full captures).
usage: python tools/step_once.py [config] [warm steps]      env: PD_B200_GRAPHS / PD_B200_OVERLAP as usual"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydreamer_b200.config import make_conf
from pydreamer_b200.dreamer import Dreamer
from pydreamer_b200.replay import synthetic_batch

cfg = sys.argv[1] if len(sys.argv) > 1 else "atari"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
conf = make_conf(cfg, device=dev)
model = Dreamer(conf).to(dev)
model.use_cuda_graph = False
model.overlap = int(os.environ.get("PD_B200_OVERLAP", "0"))
obs = synthetic_batch(conf, seed=1234, device=dev)
opts = model.init_optimizers(conf.adam_lr, conf.adam_lr_actor, conf.adam_lr_critic, conf.adam_eps)
state = model.init_state(conf.batch_size * conf.iwae_samples)


def step():
    global state
    losses, state, *_ = model.training_step(obs, state)
    for l in losses:
        l.backward()
    model.grad_clip(conf.grad_clip, conf.grad_clip_ac)
    for o in opts:
        o.step()


for _ in range(warm):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
