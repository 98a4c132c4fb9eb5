O=gpurun_out
timeout 900 python -m pytest tests/test_dreamer_gpu.py -q -x -s -k "persistent or product_arm" > $O/r3h_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|identical-trajectory" $O/r3h_tests.log | tail -5; grep -E "FAILED|^E  " $O/r3h_tests.log | head -12 | cut -c1-600
for v in 1 0; do for c in atari atari_iwae; do PD_B200_K1_GROUPED_W=$v timeout 300 python tools/k1_time.py $c > $O/r3h_k1_${c}_$v.json 2> $O/r3h_k1.err; echo "k1 $c grouped=$v rc=$? $(cat $O/r3h_k1_${c}_$v.json | cut -c1-400)"; tail -1 $O/r3h_k1.err; done; done
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3h_bench_$1.json 2> $O/r3h_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3h_bench_$1.json)"; tail -1 $O/r3h_bench_$1.err; }
b g1 PD_B200_K1_GROUPED_W=1
b g0 PD_B200_K1_GROUPED_W=0
