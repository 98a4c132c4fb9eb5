O=gpurun_out
timeout 900 python -m pytest tests/test_dreamer_gpu.py -q -x -s -k "persistent_rssm_kernel or product_arm" > $O/r2q_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|identical-trajectory" $O/r2q_tests.log | tail -5; grep -E "FAILED|^E  " $O/r2q_tests.log | head -12 | cut -c1-600
for c in atari atari_iwae; do timeout 300 python tools/k1_time.py $c > $O/r2q_k1_$c.json 2> $O/r2q_k1_$c.err; echo "k1 $c rc=$? $(cat $O/r2q_k1_$c.json | cut -c1-600)"; tail -2 $O/r2q_k1_$c.err; done
for v in 1 0; do
PD_B200_PERSISTENT_RSSM=$v timeout 900 python bench.py --config atari_iwae --steps 10 --warmup 3 --watchdog 800 --no-cpu-baseline --no-ref-gpu > $O/r2q_bench_iwae_$v.json 2> $O/r2q_bench_iwae_$v.err; echo "bench iwae persistent=$v rc=$? $(cut -c1-200 $O/r2q_bench_iwae_$v.json)"; tail -2 $O/r2q_bench_iwae_$v.err
done
timeout 600 python bench.py --config dmc --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2q_bench_dmc.json 2> $O/r2q_bench_dmc.err; echo "bench dmc rc=$? $(cut -c1-200 $O/r2q_bench_dmc.json)"
