O=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "mlp_layer" > $O/r2k_mlp_tests.log 2>&1; echo "mlp kernel tests rc=$?"; tail -3 $O/r2k_mlp_tests.log | cut -c1-400
for v in 1 0; do
PD_B200_FUSED_MLP=$v timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu --dump-gemm-profile $O/r2k_gemm_profile_$v.json > $O/r2k_bench_$v.json 2> $O/r2k_bench_$v.err; echo "bench fused=$v rc=$? $(cut -c1-200 $O/r2k_bench_$v.json)"; tail -2 $O/r2k_bench_$v.err
done
