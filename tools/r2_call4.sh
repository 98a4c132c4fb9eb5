# Round 2, call 4 (one B200): 8-warp GEMM epilogue + cheap ELU + gather/padding changes: full GPU suite (parity tables dumped),
# bench with the per-shape table, launch list.
O=gpurun_out
mkdir -p $O/parity
PD_B200_PARITY_DUMP=$O/parity timeout 1200 python -m pytest tests -m gpu -q -s > $O/r2d_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" $O/r2d_tests.log | tail -2
grep -E "^\[|free-running|FAILED|^E  " $O/r2d_tests.log | cut -c1-700 | head -40
timeout 900 python bench.py --steps 20 --warmup 5 --watchdog 800 --no-cpu-baseline --no-ref-gpu --dump-gemm-profile $O/r2d_gemm_profile.json > $O/r2d_bench.json 2> $O/r2d_bench.err; echo "bench rc=$?"; cut -c1-250 $O/r2d_bench.json; tail -3 $O/r2d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2d_launches.csv python tools/step_once.py atari 2 > $O/r2d_ncu_list.log 2>&1; echo "launch list rc=$?"
