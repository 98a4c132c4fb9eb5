"""One launch each of the implicit-convolution GEMM modes (pd_conv_gemm modes 1, 2, 3 at the decoder's 13->30 layer of the Atari
shape: 2500 images, 30x30x48 gradient image, k = 6, 96 channels) and of the fp16 2-CTA GEMM [2500, 6144, 2048], bracketed by
cudaProfilerStart/Stop after a warm-up pass: the target of `ncu --set full --profile-from-start off` captures.  Also prints
CUDA-event times of the same launches (not under the profiler when run plainly).
usage: python tools/conv_gemm_once.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydreamer_b200.ops import NativeOps

dev = "cuda:0"
ops = NativeOps(dev)
NB, H, C, k, odim = 2500, 30, 48, 6, 96
P = (H - k) // 2 + 1
pixels, K, cpad = NB * P * P, k * k * C, (C + 31) // 32 * 32
g = torch.Generator(device=dev).manual_seed(1)
X = torch.randn(NB, H, H, C, device=dev, generator=g)
Wk = torch.randn(odim, K, device=dev, generator=g) * 0.05
Ot = torch.randn(pixels, odim, device=dev, generator=g)
C1 = torch.empty(pixels, odim, device=dev)
C2 = torch.zeros(k * k * cpad, odim, device=dev)
C3 = torch.zeros(odim, k * k * cpad, device=dev)
A16 = torch.randn(2500, 2048, device=dev, generator=g).half()
B16 = torch.randn(6144, 2048, device=dev, generator=g).half()
Cf = torch.empty(2500, 6144, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timed(fn):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1000, 1)


WkT = Wk.t().contiguous()
runs = dict(conv_mode1=lambda: ops.conv_gemm(1, X, k, Wk, C1),
            conv_mode1_w_mn=lambda: ops.conv_gemm(1, X, k, WkT, C1, o_mn=True),
            conv_mode2=lambda: ops.conv_gemm(2, X, k, Ot, C2),
            conv_mode3=lambda: ops.conv_gemm(3, X, k, Ot, C3),
            gemm_f16_2cta=lambda: ops.gemm_f16(A16, B16, Cf))
for f in runs.values():
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
us = {n: timed(f) for n, f in runs.items()}
torch.cuda.cudart().cudaProfilerStop()
flops = dict(conv_mode1=2.0 * pixels * K * odim, conv_mode1_w_mn=2.0 * pixels * K * odim, conv_mode2=2.0 * pixels * K * odim, conv_mode3=2.0 * pixels * K * odim,
             gemm_f16_2cta=2.0 * 2500 * 6144 * 2048)
print(json.dumps({n: dict(us=us[n], tflops=round(flops[n] / us[n] / 1e6, 1)) for n in runs}))
