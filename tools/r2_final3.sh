O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r3y_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r3y_tests.log | cut -c1-300; grep -E "FAILED|^E  " $O/r3y_tests.log | head -8 | cut -c1-400
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r3y_bench_$1.json 2> $O/r3y_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3y_bench_$1.json)"; tail -1 $O/r3y_bench_$1.err; }
b pm1 PD_GEMM_PLAIN_M2=1
b pm0 PD_GEMM_PLAIN_M2=0
b pm1b PD_GEMM_PLAIN_M2=1
b pm0b PD_GEMM_PLAIN_M2=0
