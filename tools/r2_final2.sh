# Round-2 closing run on one B200: the whole GPU suite, smoke(), the default bench line with both baselines.
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r3z_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r3z_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r3z_smoke.log
timeout 900 python bench.py --watchdog 800 --dump-gemm-profile $O/r3z_gemm_profile_by_shape.json > $O/r3z_bench_ours.json 2> $O/r3z_bench_ours.err; echo "bench rc=$?"; cut -c1-300 $O/r3z_bench_ours.json; tail -2 $O/r3z_bench_ours.err
timeout 600 python bench.py --config dmc --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3z_bench_dmc.json 2> $O/r3z_bench_dmc.err; echo "dmc rc=$? $(cut -c1-200 $O/r3z_bench_dmc.json)"
timeout 900 python bench.py --config atari_iwae --steps 10 --warmup 3 --watchdog 800 --no-cpu-baseline --no-ref-gpu > $O/r3z_bench_iwae.json 2> $O/r3z_bench_iwae.err; echo "iwae rc=$? $(cut -c1-200 $O/r3z_bench_iwae.json)"
