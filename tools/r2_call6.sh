# Round 2, call 6 (one B200): first run of the third-generation persistent forward kernel (pd_rssm_fwd3.cu) + BPTT A/B.
O=gpurun_out
timeout 600 python -m pytest tests/test_dreamer_gpu.py -q -s -x -k "persistent or product_arm or full_atari" > $O/r2f_k1_tests.log 2>&1; echo "k1 tests rc=$?"; grep -E "passed|failed|error" $O/r2f_k1_tests.log | tail -2; grep -E "^persistent|full-size|Error|error|assert|FAILED|Warning" $O/r2f_k1_tests.log | cut -c1-500 | head -20
timeout 200 python tools/k1_time.py atari > $O/r2f_k1_time.json 2> $O/r2f_k1_time.err; echo "k1 time rc=$?"; cat $O/r2f_k1_time.json; tail -3 $O/r2f_k1_time.err
timeout 200 python tools/k1_time.py dmc > $O/r2f_k1_time_dmc.json 2> $O/r2f_k1_time_dmc.err; cat $O/r2f_k1_time_dmc.json
for v in 0 1; do
PD_B200_PERSISTENT_BPTT=$v timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2f_bench_bptt$v.json 2> $O/r2f_bench_bptt$v.err; echo "bench bptt=$v rc=$? $(cut -c1-200 $O/r2f_bench_bptt$v.json)"; tail -2 $O/r2f_bench_bptt$v.err
done
timeout 900 python -m pytest tests -m gpu -q > $O/r2f_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2f_tests.log | tail -2; grep -E "FAILED" $O/r2f_tests.log | head
