# Round-end check on one B200: GPU test suite, smoke(), default bench (ours), reference-arm bench.
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --watchdog 500 > gpurun_out/final_bench_ours.json 2> gpurun_out/final_bench_ours.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench_ours.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/final_bench_ref.json
