O=gpurun_out
timeout 900 python -m pytest tests/test_dreamer_gpu.py -q -x -s -k "persistent_rssm_kernel or product_arm" > $O/r2r_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|identical-trajectory" $O/r2r_tests.log | tail -5; grep -E "FAILED|^E  " $O/r2r_tests.log | head -12 | cut -c1-600
for c in atari atari_iwae; do timeout 300 python tools/k1_time.py $c > $O/r2r_k1_$c.json 2> $O/r2r_k1_$c.err; echo "k1 $c rc=$? $(cat $O/r2r_k1_$c.json | cut -c1-600)"; tail -2 $O/r2r_k1_$c.err; done
timeout 300 python tools/conv_gemm_once.py > $O/r2r_conv_gemm_once.json 2> $O/r2r_conv_gemm_once.err; echo "conv once rc=$? $(cat $O/r2r_conv_gemm_once.json)"; tail -2 $O/r2r_conv_gemm_once.err
cap() { timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -c $3 -f -o $O/r2r_$1 ${@:4} > $O/r2r_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"; }
cap k1fwd_final rssm_unroll_fwd3 1 python tools/step_once.py atari 2
PD_B200_PERSISTENT_BPTT=1 cap k1bwd_final rssm_unroll_bwd 1 python tools/step_once.py atari 2
cap conv_modes_and_f16 pd_gemm 4 python tools/conv_gemm_once.py
ls -la $O/*.ncu-rep | tail -5
