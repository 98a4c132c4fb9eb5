# Round 2, call 5 (one B200): first run of the persistent BPTT kernel.
O=gpurun_out
timeout 600 python -m pytest tests/test_dreamer_gpu.py -q -s -x -k "persistent_bptt or product_arm" > $O/r2e_k1b_tests.log 2>&1; echo "k1b tests rc=$?"; grep -E "passed|failed|error" $O/r2e_k1b_tests.log | tail -2; grep -E "^\[|Error|error|assert|FAILED|Warning" $O/r2e_k1b_tests.log | cut -c1-600 | head -30
timeout 200 python tools/k1b_time.py > $O/r2e_k1b_time.json 2> $O/r2e_k1b_time.err; echo "k1b time rc=$?"; cat $O/r2e_k1b_time.json; tail -3 $O/r2e_k1b_time.err
timeout 900 python -m pytest tests -m gpu -q -s > $O/r2e_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2e_tests.log | tail -2; grep -E "^\[atari|^\[dmc|free-running|FAILED|^E  " $O/r2e_tests.log | cut -c1-500 | head -30
timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2e_bench.json 2> $O/r2e_bench.err; echo "bench rc=$?"; cut -c1-250 $O/r2e_bench.json; tail -3 $O/r2e_bench.err
PD_B200_PERSISTENT_BPTT=0 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2e_bench_chain.json 2> $O/r2e_bench_chain.err; echo "bench chain rc=$?"; cut -c1-250 $O/r2e_bench_chain.json
