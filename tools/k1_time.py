"""Times pd_rssm_unroll_fwd (persistent posterior unroll, Atari shape by default): CUDA events around the call."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydreamer_b200.config import make_conf
from pydreamer_b200.dreamer import Dreamer
from pydreamer_b200.replay import synthetic_batch

conf = make_conf(sys.argv[1] if len(sys.argv) > 1 else "atari", device="cuda:0")
obs = synthetic_batch(conf, seed=1, device="cuda:0")
model = Dreamer(conf).to("cuda:0")
model.persistent_rssm = True
model.overlap = 0
state = model.init_state(conf.batch_size)
with torch.no_grad():
    model.training_step(obs, state)
ops = model.ops
orig, times = ops.rssm_unroll_fwd, []
def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(*a, **k); e1.record(); times.append((e0, e1))
ops.rssm_unroll_fwd = timed
with torch.no_grad():
    for _ in range(3):
        model.training_step(obs, state)
torch.cuda.synchronize()
print(json.dumps(dict(kernel_ms=[round(a.elapsed_time(b), 3) for a, b in times])))
