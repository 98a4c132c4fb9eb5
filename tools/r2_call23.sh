O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv" > $O/r2v_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r2v_kernel_tests.log | cut -c1-600
timeout 300 python tools/conv_gemm_once.py > $O/r2v_conv_once_default.json 2> $O/r2v_conv_once.err; echo "conv once default rc=$? $(cat $O/r2v_conv_once_default.json)"; tail -1 $O/r2v_conv_once.err
PD_GEMM_TRIM_B=0 PD_GEMM_MN3=0 timeout 300 python tools/conv_gemm_once.py > $O/r2v_conv_once_off.json 2> $O/r2v_conv_once.err; echo "conv once trim=0 mn3=0 rc=$? $(cat $O/r2v_conv_once_off.json)"; tail -1 $O/r2v_conv_once.err
PD_GEMM_2CTA_MINM=384 timeout 300 python tools/gemm_bench.py 400,3072,37500,1,1,1 400,400,37500,1,1,1 400,400,2500,1,1,1 400,3072,2500,1,1,1 > $O/r2v_gemm_m384.jsonl 2>&1; cat $O/r2v_gemm_m384.jsonl | cut -c1-200
timeout 300 python tools/gemm_bench.py 400,3072,37500,1,1,1 400,400,37500,1,1,1 400,400,2500,1,1,1 400,3072,2500,1,1,1 > $O/r2v_gemm_m512.jsonl 2>&1; cat $O/r2v_gemm_m512.jsonl | cut -c1-200
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2v_bench_$1.json 2> $O/r2v_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r2v_bench_$1.json)"; tail -1 $O/r2v_bench_$1.err; }
b default A=1
b minm384 PD_GEMM_2CTA_MINM=384
b trim0 PD_GEMM_TRIM_B=0
b default2 A=1
