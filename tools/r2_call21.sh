O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv" > $O/r2t_kernel_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 $O/r2t_kernel_tests.log | cut -c1-400
for v in 1 0; do PD_GEMM_MN3=$v timeout 300 python tools/conv_gemm_once.py > $O/r2t_conv_once_$v.json 2> $O/r2t_conv_once_$v.err; echo "conv once mn3=$v rc=$? $(cat $O/r2t_conv_once_$v.json)"; done
PD_GEMM_MN3=1 timeout 300 python tools/gemm_bench.py 400,3072,37500,1,1,1 400,400,37500,1,1,1 108,48,2250000,1,1,1 6144,2048,2500,1,1,1 2500,3072,400,0,1 37500,400,400,0,1 > $O/r2t_gemm_mn3_1.jsonl 2>&1; cat $O/r2t_gemm_mn3_1.jsonl | cut -c1-300
PD_GEMM_MN3=0 timeout 300 python tools/gemm_bench.py 400,3072,37500,1,1,1 400,400,37500,1,1,1 108,48,2250000,1,1,1 6144,2048,2500,1,1,1 2500,3072,400,0,1 37500,400,400,0,1 > $O/r2t_gemm_mn3_0.jsonl 2>&1; cat $O/r2t_gemm_mn3_0.jsonl | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2t_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2t_tests.log | tail -2; grep -E "FAILED|^E  " $O/r2t_tests.log | head -12 | cut -c1-600
for v in 1 0; do
PD_GEMM_MN3=$v timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2t_bench_$v.json 2> $O/r2t_bench_$v.err; echo "bench mn3=$v rc=$? $(cut -c1-200 $O/r2t_bench_$v.json)"; tail -2 $O/r2t_bench_$v.err
done
