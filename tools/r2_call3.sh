# Round 2, call 3 (one B200): full GPU suite incl. the new full-size element-wise parity tests, the new bench line,
# ncu of the conv1 GEMM (first tcgen05 launch of a step) and of the reworked col2im_imgloss.
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 > $O/r2c_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $O/r2c_tests.log | tail -3
grep -E "^\[|free-running|full-size|FAILED|Error|assert" $O/r2c_tests.log | cut -c1-900 | head -60
timeout 900 python bench.py --steps 20 --warmup 5 --watchdog 800 --dump-gemm-profile $O/r2c_gemm_profile.json > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench rc=$?"; cut -c1-250 $O/r2c_bench.json; tail -3 $O/r2c_bench.err
cap() { timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -c $3 -f -o $O/r2c_$1 python tools/step_once.py atari 2 > $O/r2c_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"; }
cap conv1gemm 'pd_gemm_tf32_kernel' 1
cap imgloss col2im_imgloss 1
cap col2imv4 col2im_v4 3
cap biasact bias_act_bwd 2
ls -la $O | grep r2c_ | awk '{print $5, $9}'
