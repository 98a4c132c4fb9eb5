"""Micro-benchmark of pd_gemm on individual shapes (CUDA events; also the target of ncu --set full captures).
usage: python tools/gemm_bench.py M,N,K[,a_mn,b_mn,acc,bias,res] ... [--reps 20]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydreamer_b200.ops import NativeOps

def main():
    reps = 20
    shapes = []
    args = [a for a in sys.argv[1:] if a not in ('--chain', '--f16')]
    if "--reps" in args:
        i = args.index("--reps"); reps = int(args[i + 1]); del args[i:i + 2]
    for a in args:
        v = [int(x) for x in a.split(",")]
        v += [0] * (8 - len(v))
        shapes.append(v)
    ops = NativeOps("cuda:0")
    out = []
    if "--chain" in sys.argv:
        # dependent chain of identical small GEMMs replayed from a CUDA graph: per-launch latency inside a step
        args2 = [a for a in sys.argv[1:] if not a.startswith("--") and "," in a]
        for a in args2:
            v = [int(x) for x in a.split(",")] ; v += [0] * (8 - len(v))
            M, N, K, a_mn, b_mn, acc, bias, res = v
            A = torch.randn((M, K), device="cuda"); B = torch.randn((N, K), device="cuda")
            Cs = [torch.zeros(M, N, device="cuda") for _ in range(2)]
            bv = torch.randn(N, device="cuda") if bias else None
            rv = torch.randn(M, N, device="cuda") if res else None
            n = 200
            def body():
                for i in range(n):
                    ops.gemm(A, B, Cs[i & 1], bias=bv, res=rv)
            body(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            print(json.dumps(dict(chain=[M, N, K, bias, res], us_per_gemm=1000 * e0.elapsed_time(e1) / n,
                                  maxsplit=os.environ.get("PD_GEMM_SKINNY_MAXSPLIT", ""))))
        return
    for (M, N, K, a_mn, b_mn, acc, bias, res) in shapes:
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        bv = torch.randn(N, device="cuda") if bias else None
        rv = torch.randn(M, N, device="cuda") if res else None
        flush = torch.empty(64 * 1024 * 1024, device="cuda")
        f16 = "--f16" in sys.argv
        if f16:
            A16, B16 = A.half(), B.half()
        def run():
            if f16:
                ops.gemm_f16(A16, B16, C, bias=bv, res=rv)
            else:
                ops.gemm(A, B, C, a_mn=bool(a_mn), b_mn=bool(b_mn), accumulate=bool(acc), bias=bv, res=rv)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()                     # 256 MB write: evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        fl = 2.0 * M * N * K
        by = 4.0 * (M * K + N * K + M * N)
        out.append(dict(shape=[M, N, K, a_mn, b_mn, acc, bias, res], ms=med, tflops=fl / med / 1e9, gbs=by / med / 1e6))
        print(json.dumps(out[-1]))

if __name__ == "__main__":
    main()
