# Round 2, call 1 (one B200): sanity of the round-1 state, default bench, reference eager on the same GPU, the launch list of
# one steady-state step, --set full captures of the kernels VERDICT r01 asked for, and the two never-run opt-in kernels.
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2a_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --watchdog 300 --dump-gemm-profile $O/r2a_gemm_profile.json > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r2a_bench.json
timeout 300 python bench.py --impl reference --ref-device cuda --steps 20 --warmup 5 > $O/r2a_ref_gpu.json 2> $O/r2a_ref_gpu.err; echo "ref gpu rc=$?"; cut -c1-300 $O/r2a_ref_gpu.json
# launch list of one eager single-stream step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2a_launches.csv python tools/step_once.py atari 2 > $O/r2a_ncu_list.log 2>&1; echo "launch list rc=$?"
# full captures (one launch each unless stated)
cap() { # name regex count
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -c $3 -f -o $O/r2a_$1 python tools/step_once.py atari 2 > $O/r2a_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap k1fwd rssm_unroll_fwd 1
cap imgloss col2im_imgloss 1
cap col2im "col2im_kernel" 3
cap im2col im2col 2
cap lnelu ln_elu_fwd 3
cap adamw adamw 1
cap gemm2cta pd_gemm_tf32_2cta 6
cap grubwd gru_bwd 1
# opt-in kernels that never ran on a GPU
PD_B200_K1_STAGING=tma timeout 120 python tools/k1_time.py > $O/r2a_k1_time_tma.json 2> $O/r2a_k1_time_tma.err; echo "k1 tma rc=$?"; cat $O/r2a_k1_time_tma.json; tail -3 $O/r2a_k1_time_tma.err
timeout 120 python tools/k1_time.py > $O/r2a_k1_time.json 2> $O/r2a_k1_time.err; echo "k1 rc=$?"; cat $O/r2a_k1_time.json
PD_B200_K1_STAGING=tma timeout 300 python -m pytest tests/test_dreamer_gpu.py -x -q -k "persistent" > $O/r2a_k1_tests_tma.log 2>&1; echo "tma tests rc=$?"; tail -3 $O/r2a_k1_tests_tma.log
PD_B200_TEST_EXPERIMENTAL=1 PD_B200_DIRECT_CONV1=1 timeout 300 python -m pytest tests -m gpu -x -q -k "conv1_direct" > $O/r2a_conv1_tests.log 2>&1; echo "conv1 tests rc=$?"; tail -3 $O/r2a_conv1_tests.log
PD_B200_DIRECT_CONV1=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --watchdog 200 > $O/r2a_bench_conv1.json 2> $O/r2a_bench_conv1.err; echo "bench conv1 rc=$?"; cut -c1-200 $O/r2a_bench_conv1.json
ls -la $O | grep r2a_ | awk '{print $5, $9}'
