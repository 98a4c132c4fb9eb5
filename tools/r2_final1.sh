# Round-2 final single-GPU record: smoke(), default bench line (with cpu_baseline and reference_gpu_eager), the other two
# single-GPU configurations, the reference arm.
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2y_smoke.log
timeout 900 python bench.py --watchdog 800 --dump-gemm-profile $O/r2y_gemm_profile_by_shape.json > $O/r2y_bench_ours.json 2> $O/r2y_bench_ours.err; echo "bench rc=$?"; cut -c1-300 $O/r2y_bench_ours.json; tail -2 $O/r2y_bench_ours.err
timeout 600 python bench.py --config dmc --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2y_bench_dmc.json 2> $O/r2y_bench_dmc.err; echo "dmc rc=$? $(cut -c1-200 $O/r2y_bench_dmc.json)"
timeout 900 python bench.py --config atari_iwae --steps 10 --warmup 3 --watchdog 800 --no-cpu-baseline --no-ref-gpu > $O/r2y_bench_iwae.json 2> $O/r2y_bench_iwae.err; echo "iwae rc=$? $(cut -c1-200 $O/r2y_bench_iwae.json)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2y_bench_ref.json 2> $O/r2y_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2y_bench_ref.json
