timeout 400 python -m pytest tests/test_dreamer_gpu.py -x -q -k "persistent" > gpurun_out/k1_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/k1_tests.log
for v in 1 0; do
PD_B200_PERSISTENT_RSSM=$v timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --watchdog 100 > gpurun_out/k1_bench_$v.json 2> gpurun_out/k1_bench_$v.err; echo "bench K1=$v rc=$? $(cut -c1-150 gpurun_out/k1_bench_$v.json)"
done
