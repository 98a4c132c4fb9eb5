O=gpurun_out
timeout 200 python tools/k1_time.py atari > $O/r2g_k1_time.json 2> $O/r2g_k1_time.err; echo "k1 time rc=$?"; cat $O/r2g_k1_time.json; tail -3 $O/r2g_k1_time.err
timeout 200 python tools/k1b_time.py atari > $O/r2g_k1b_time.json 2> $O/r2g_k1b_time.err; echo "k1b time rc=$?"; cat $O/r2g_k1b_time.json; tail -3 $O/r2g_k1b_time.err
timeout 300 python -m pytest tests/test_dreamer_gpu.py -q -x -k "persistent" > $O/r2g_k1_tests.log 2>&1; echo "k1 tests rc=$?"; tail -2 $O/r2g_k1_tests.log
