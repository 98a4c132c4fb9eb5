O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "fused or gemm or conv or col2im" > $O/r2x_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r2x_kernel_tests.log | cut -c1-800; grep -E "^E  " $O/r2x_kernel_tests.log | head -8 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2x_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2x_tests.log | tail -2; grep -E "FAILED|^E  " $O/r2x_tests.log | head -12 | cut -c1-600
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2x_bench_$1.json 2> $O/r2x_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r2x_bench_$1.json)"; tail -1 $O/r2x_bench_$1.err; }
b fuse1 PD_B200_FUSE_ACTBWD=1
b fuse0 PD_B200_FUSE_ACTBWD=0
b fuse1b PD_B200_FUSE_ACTBWD=1
b fuse0b PD_B200_FUSE_ACTBWD=0
