O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r3a_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r3a_tests.log | tail -2; grep -E "FAILED|^E  " $O/r3a_tests.log | head -12 | cut -c1-600
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu --dump-gemm-profile $O/r3a_gemm_profile_$1.json > $O/r3a_bench_$1.json 2> $O/r3a_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r3a_bench_$1.json)"; tail -1 $O/r3a_bench_$1.err; }
b a A=1
b b A=1
