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
state = model.init_state(conf.batch_size * conf.iwae_samples)
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
bar = model._buf("k1.bar", 16, dtype=torch.int32)
ns = bar[2:16].view(torch.int64).tolist()
names = ["prologue+gh0", "A_gather_ln", "B_gi_gru", "C_hh_ph_partials", "Cp_ln", "D_logits_sample"]
T = conf.batch_length
print(json.dumps(dict(kernel_ms=[round(a.elapsed_time(b), 3) for a, b in times],
                      phase_us_per_step={n: round(v / 1000 / (1 if i == 0 else T), 2) for i, (n, v) in enumerate(zip(names, ns))},
                      total_ms_from_phases=round(sum(ns[:6]) / 1e6, 3))))
