"""Bring-up probe: does a TMA im2col-mode load reproduce rows of the explicit im2col matrix?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydreamer_b200.ops import NativeOps

ops = NativeOps("cuda:0")
lib = ops.lib
lib.pd_dbg_im2col_load.restype = ctypes.c_int
lib.pd_dbg_im2col_load.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 14 + [ctypes.c_void_p, ctypes.c_void_p]
NB, H, W, C, k = 3, 14, 14, 96, 4
P = Q = (H - k) // 2 + 1
x = torch.randn(NB, H, W, C, device="cuda")
pat = x.unfold(1, k, 2).unfold(2, k, 2)            # (NB,P,Q,C,kh,kw)
for (lower, upper) in [(0, -(k - 1))]:
    for (m0, kh, kw, c0) in [(0, 0, 0, 0), (0, 1, 2, 32), (5, 3, 3, 64), (30, 2, 1, 0), (100, 0, 3, 32)]:
        pixels = 64
        n0, rem = divmod(m0, P * Q); p0, q0 = divmod(rem, Q)
        out = torch.full((pixels, 32), float("nan"), device="cuda")
        rc = lib.pd_dbg_im2col_load(ops.h, x.data_ptr(), NB, H, W, C, k, lower, upper, c0, 2 * q0, 2 * p0, n0, kw, kh, pixels,
                                    out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = torch.zeros(pixels, 32, device="cuda")
        flat = pat[..., kh, kw].reshape(NB * P * Q, C)
        nrow = min(pixels, NB * P * Q - m0)
        ref[:nrow] = flat[m0:m0 + nrow, c0:c0 + 32]
        ok = torch.equal(out[:nrow], ref[:nrow])
        print(f"corners ({lower},{upper}) m0={m0} tap=({kh},{kw}) c0={c0} rc={rc} match={ok} tail_zero={bool((out[nrow:] == 0).all()) if nrow < pixels else None}")
        if not ok:
            bad = (out[:nrow] != ref[:nrow]).any(1).nonzero().flatten()[:8].tolist()
            print("  first bad rows", bad, "out[0,:4]", out[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())
