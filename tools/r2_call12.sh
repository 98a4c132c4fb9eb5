O=gpurun_out
for v in 0 1; do
PD_B200_HP_MAIN=$v timeout 600 python bench.py --steps 30 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2l_bench_hp$v.json 2> $O/r2l_bench_hp$v.err; echo "bench hp=$v rc=$? $(cut -c1-200 $O/r2l_bench_hp$v.json)"; tail -2 $O/r2l_bench_hp$v.err
done
PD_B200_HP_MAIN=1 timeout 400 python -m pytest tests/test_dreamer_gpu.py -q -x -k "product_arm or optimizer_step" > $O/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2l_tests.log
