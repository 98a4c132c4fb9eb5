O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > $O/r3i_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/r3i_kernel_tests.log | cut -c1-400; grep -E "^E  " $O/r3i_kernel_tests.log | head -5 | cut -c1-300
for v in 1 0; do PD_GEMM_PLAIN_M2=$v timeout 300 python tools/gemm_bench.py 2402500,48,48,0,0,0,1 2250000,108,48 2250000,48,108,0,1 --reps 5 > $O/r3i_m2p_$v.jsonl 2>&1; echo "plain m2=$v"; cut -c1-130 $O/r3i_m2p_$v.jsonl; done
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3i_bench_$1.json 2> $O/r3i_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3i_bench_$1.json)"; tail -1 $O/r3i_bench_$1.err; }
b pm2on PD_GEMM_PLAIN_M2=1
b pm2off PD_GEMM_PLAIN_M2=0
b cols16 PD_B200_FP16_COLS=1
