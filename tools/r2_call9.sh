O=gpurun_out
timeout 400 python -m pytest tests/test_dreamer_gpu.py -q -x -k "persistent or product_arm or full_atari" > $O/r2i_k1_tests.log 2>&1; echo "k1 tests rc=$?"; tail -2 $O/r2i_k1_tests.log
timeout 200 python tools/k1_time.py atari > $O/r2i_k1_time.json 2> $O/r2i_k1_time.err; echo "k1 time rc=$?"; cat $O/r2i_k1_time.json; tail -3 $O/r2i_k1_time.err
timeout 200 python tools/k1b_time.py atari > $O/r2i_k1b_time.json 2> $O/r2i_k1b_time.err; cat $O/r2i_k1b_time.json
timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2i_bench.json 2> $O/r2i_bench.err; echo "bench rc=$? $(cut -c1-200 $O/r2i_bench.json)"; tail -2 $O/r2i_bench.err
