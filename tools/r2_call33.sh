O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv" > $O/r3e_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r3e_kernel_tests.log | cut -c1-800; grep -E "^E  " $O/r3e_kernel_tests.log | head -8 | cut -c1-400
S="2500,6144,2048 2500,6144,1000 2500,1000,2048 2500,1024,1000 2500,1000,1024 2500,400,3072 40000,400,3072 40000,400,400 2500,400,400"
for v in 1 0; do PD_GEMM_2CTA_K2=$v timeout 300 python tools/gemm_bench.py $S --f16 --reps 10 > $O/r3e_f16_k2_$v.jsonl 2>&1; echo "f16 k2=$v"; cut -c1-140 $O/r3e_f16_k2_$v.jsonl; done
for v in 1 0; do PD_GEMM_2CTA_K2=$v timeout 300 python tools/gemm_bench.py 40000,400,400 37500,400,400 2500,3072,400 40000,400,3072 --reps 10 > $O/r3e_tf32_k2_$v.jsonl 2>&1; echo "tf32 k2=$v"; cut -c1-140 $O/r3e_tf32_k2_$v.jsonl; done
timeout 900 python -m pytest tests -m gpu -q -x > $O/r3e_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r3e_tests.log | tail -2; grep -E "FAILED|^E  " $O/r3e_tests.log | head -12 | cut -c1-600
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3e_bench_$1.json 2> $O/r3e_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3e_bench_$1.json)"; tail -1 $O/r3e_bench_$1.err; }
b k2on PD_GEMM_2CTA_K2=1
b k2off PD_GEMM_2CTA_K2=0
b k2onb PD_GEMM_2CTA_K2=1
