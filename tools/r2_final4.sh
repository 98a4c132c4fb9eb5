O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r3x_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r3x_tests.log | cut -c1-300; grep -E "FAILED|^E  " $O/r3x_tests.log | head -8 | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r3x_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --watchdog 250 --no-cpu-baseline --no-ref-gpu > $O/r3x_bench.json 2> $O/r3x_bench.err; echo "bench rc=$? $(cut -c1-200 $O/r3x_bench.json)"
