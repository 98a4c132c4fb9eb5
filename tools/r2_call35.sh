O=gpurun_out
b() { timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3g_bench_$1.json 2> $O/r3g_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3g_bench_$1.json)"; tail -1 $O/r3g_bench_$1.err; }
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > $O/r3g_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/r3g_kernel_tests.log | cut -c1-400
b st4_a
cp pydreamer_b200/libpd_b200.so /tmp/new.so; cp ab_old/libpd_b200.so pydreamer_b200/libpd_b200.so
b st5_a
cp /tmp/new.so pydreamer_b200/libpd_b200.so
b st4_b
cp ab_old/libpd_b200.so pydreamer_b200/libpd_b200.so
b st5_b
cp /tmp/new.so pydreamer_b200/libpd_b200.so
