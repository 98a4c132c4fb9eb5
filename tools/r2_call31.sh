O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused or gemm or conv or col2im" > $O/r3c_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r3c_kernel_tests.log | cut -c1-800; grep -E "^E  " $O/r3c_kernel_tests.log | head -8 | cut -c1-400
for v in 1 0; do PD_GEMM_CONV_M2=$v timeout 300 python tools/conv_gemm_once.py > $O/r3c_conv_once_$v.json 2> $O/r3c_conv_once_$v.err; echo "conv once m2=$v rc=$? $(cat $O/r3c_conv_once_$v.json)"; tail -1 $O/r3c_conv_once_$v.err; done
timeout 900 python -m pytest tests -m gpu -q -x > $O/r3c_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r3c_tests.log | tail -2; grep -E "FAILED|^E  " $O/r3c_tests.log | head -12 | cut -c1-600
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3c_bench_$1.json 2> $O/r3c_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3c_bench_$1.json)"; tail -1 $O/r3c_bench_$1.err; }
b m2on PD_GEMM_CONV_M2=1
b m2off PD_GEMM_CONV_M2=0
b m2onb PD_GEMM_CONV_M2=1
