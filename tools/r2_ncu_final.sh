# ncu --set full captures of the kernels as shipped at the end of round 2
O=gpurun_out
cap() { timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -c $3 -f -o $O/r3z_$1 ${@:4} > $O/r3z_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"; }
cap conv_modes_final pd_gemm 5 python tools/conv_gemm_once.py
cap k1fwd_grouped rssm_unroll_fwd3 1 python tools/step_once.py atari 2
ls -la $O/r3z_*.ncu-rep
