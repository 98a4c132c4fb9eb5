"""Times pd_rssm_unroll_bwd (persistent BPTT, Atari shape by default) with CUDA events around the call, next to the
per-timestep chain it replaces (same step, PD_B200_PERSISTENT_BPTT toggled)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydreamer_b200.config import make_conf
from pydreamer_b200.dreamer import Dreamer
from pydreamer_b200.replay import synthetic_batch

cfg = sys.argv[1] if len(sys.argv) > 1 else "atari"
conf = make_conf(cfg, device="cuda:0")
obs = synthetic_batch(conf, seed=1, device="cuda:0")
out = {}
for mode in (True, False):
    model = Dreamer(conf).to("cuda:0")
    model.persistent_bptt = mode
    model.overlap = 0
    model.use_cuda_graph = False
    state = model.init_state(conf.batch_size * conf.iwae_samples)
    step = lambda: [l.backward() for l in model.training_step(obs, state)[0]]
    step(); torch.cuda.synchronize()
    ops = model.ops
    times = []
    if mode:
        orig = ops.rssm_unroll_bwd
        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); orig(*a, **k); e1.record(); times.append((e0, e1))
        ops.rssm_unroll_bwd = timed
    else:
        # the chain: everything between the first cat_st_bwd and the last per-step z_mlp gemm of _wm_backward
        orig_c = ops.cat_st_bwd
        st = {"first": None, "n": 0}
        T = conf.batch_length
        def c_timed(*a, **k):
            if st["n"] % T == 0:
                st["first"] = torch.cuda.Event(enable_timing=True); st["first"].record()
            orig_c(*a, **k); st["n"] += 1
        ops.cat_st_bwd = c_timed
        orig_g = ops.colsum
        def g_timed(*a, **k):
            if st["first"] is not None:
                e1 = torch.cuda.Event(enable_timing=True); e1.record(); times.append((st["first"], e1)); st["first"] = None
            orig_g(*a, **k)
        ops.colsum = g_timed
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    out["persistent_ms" if mode else "chain_ms"] = [round(a.elapsed_time(b), 3) for a, b in times]
    if mode:
        T = conf.batch_length
        ns = model._buf("k1b.bar", 16, dtype=torch.int32)[2:16].view(torch.int64).tolist()
        names = ["P1_dpost", "P2_dpin", "P3_ln2_bwd", "P4_dh_gru_bwd", "P67_dhin_dza", "P8_ln1_bwd", "P9_dz"]
        out["phase_us_per_step"] = {n: round(v / 1000 / T, 2) for n, v in zip(names, ns)}
    del model
print(json.dumps(out))
