O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2w_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2w_tests.log | tail -2; grep -E "FAILED|^E  " $O/r2w_tests.log | head -12 | cut -c1-600
b() { env $2 timeout 600 python bench.py --steps 20 --warmup 5 --watchdog 500 --no-cpu-baseline --no-ref-gpu > $O/r2w_bench_$1.json 2> $O/r2w_bench_$1.err; echo "bench $1 rc=$? $(cut -c1-200 $O/r2w_bench_$1.json)"; tail -1 $O/r2w_bench_$1.err; }
b ov3 PD_B200_OVERLAP=3
b ov7 PD_B200_OVERLAP=7
b ov3b PD_B200_OVERLAP=3
b ov7b PD_B200_OVERLAP=7
cap() { timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -c $3 -f -o $O/r2w_$1 ${@:4} > $O/r2w_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"; }
cap conv_modes_k64 pd_gemm 5 python tools/conv_gemm_once.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2w_launches.csv python tools/step_once.py atari 2 > $O/r2w_ncu_list.log 2>&1; echo "launch list rc=$?"
gzip -f $O/r2w_launches.csv
