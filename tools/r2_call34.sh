O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > $O/r3f_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/r3f_kernel_tests.log | cut -c1-400
S="2500,6144,2048 2500,6144,1000 2500,1000,2048 40000,400,3072 40000,400,400"
for v in 1 0; do PD_GEMM_2CTA_K2=$v timeout 300 python tools/gemm_bench.py $S --f16 --reps 10 > $O/r3f_f16_k2_$v.jsonl 2>&1; echo "f16 k2=$v"; cut -c1-120 $O/r3f_f16_k2_$v.jsonl; done
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3f_bench_$1.json 2> $O/r3f_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3f_bench_$1.json)"; tail -1 $O/r3f_bench_$1.err; }
b k2on PD_GEMM_2CTA_K2=1
b k2off PD_GEMM_2CTA_K2=0
b k2onb PD_GEMM_2CTA_K2=1
b k2offb PD_GEMM_2CTA_K2=0
