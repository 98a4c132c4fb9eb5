O=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "fp16_column or conv_data_movement or gemm" > $O/r2n_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r2n_kernel_tests.log | cut -c1-400
mkdir -p $O/parity3
PD_B200_PARITY_DUMP=$O/parity3 timeout 900 python -m pytest tests -m gpu -q -s > $O/r2n_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2n_tests.log | tail -2; grep -E "FAILED|^E  |forward tensors" $O/r2n_tests.log | head -12 | cut -c1-600
for v in 1 0; do
PD_B200_FP16_COLS=$v timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2n_bench_$v.json 2> $O/r2n_bench_$v.err; echo "bench fp16cols=$v rc=$? $(cut -c1-200 $O/r2n_bench_$v.json)"; tail -2 $O/r2n_bench_$v.err
done
