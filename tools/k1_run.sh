# Persistent-RSSM A/B on one B200: module tests of both unroll paths, phase timers, then bench.py with the kernel on/off
# and (third run) with the TMA-tile staging variant.
timeout 400 python -m pytest tests/test_dreamer_gpu.py -x -q -k "persistent" > gpurun_out/k1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/k1_tests.log
timeout 120 python tools/k1_time.py > gpurun_out/k1_time.json 2> gpurun_out/k1_time.err; cat gpurun_out/k1_time.json
PD_B200_K1_STAGING=tma timeout 120 python tools/k1_time.py > gpurun_out/k1_time_tma.json 2> gpurun_out/k1_time_tma.err; cat gpurun_out/k1_time_tma.json
PD_B200_K1_STAGING=tma timeout 400 python -m pytest tests/test_dreamer_gpu.py -x -q -k "persistent" > gpurun_out/k1_tests_tma.log 2>&1; echo "tma tests rc=$?"; tail -3 gpurun_out/k1_tests_tma.log
for v in 1 0; do
PD_B200_PERSISTENT_RSSM=$v timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --watchdog 100 > gpurun_out/k1_bench_$v.json 2> gpurun_out/k1_bench_$v.err; echo "bench K1=$v rc=$? $(cut -c1-150 gpurun_out/k1_bench_$v.json)"
done
