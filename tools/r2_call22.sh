O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv" > $O/r2u_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r2u_kernel_tests.log | cut -c1-600
for v in 1 0; do PD_GEMM_CONV_K64=$v timeout 300 python tools/conv_gemm_once.py > $O/r2u_conv_once_$v.json 2> $O/r2u_conv_once_$v.err; echo "conv once k64=$v rc=$? $(cat $O/r2u_conv_once_$v.json)"; tail -1 $O/r2u_conv_once_$v.err; done
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2u_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2u_tests.log | tail -2; grep -E "FAILED|^E  " $O/r2u_tests.log | head -12 | cut -c1-600
for v in 1 0 1; do
PD_GEMM_CONV_K64=$v timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2u_bench_$v.json 2> $O/r2u_bench_$v.err; echo "bench k64=$v rc=$? $(cut -c1-200 $O/r2u_bench_$v.json)"; tail -2 $O/r2u_bench_$v.err
done
