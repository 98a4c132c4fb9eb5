O=gpurun_out
timeout 400 python -m pytest tests/test_dreamer_gpu.py -q -x -s -k "persistent or product_arm or full_atari" > $O/r2h_k1_tests.log 2>&1; echo "k1 tests rc=$?"; grep -E "passed|failed|error" $O/r2h_k1_tests.log | tail -2; grep -E "^persistent|full-size|Error|error|assert|FAILED" $O/r2h_k1_tests.log | cut -c1-400 | head -12
timeout 200 python tools/k1_time.py atari > $O/r2h_k1_time.json 2> $O/r2h_k1_time.err; echo "k1 time rc=$?"; cat $O/r2h_k1_time.json; tail -3 $O/r2h_k1_time.err
timeout 200 python tools/k1_time.py dmc > $O/r2h_k1_time_dmc.json 2> $O/r2h_k1_time_dmc.err; cat $O/r2h_k1_time_dmc.json
timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2h_bench.json 2> $O/r2h_bench.err; echo "bench rc=$? $(cut -c1-200 $O/r2h_bench.json)"; tail -2 $O/r2h_bench.err
timeout 900 python -m pytest tests -m gpu -q > $O/r2h_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2h_tests.log | tail -2; grep -E "FAILED" $O/r2h_tests.log | head
