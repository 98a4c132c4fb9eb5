O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pd_gemm_tf32_kernel -c 1 -f -o $O/r2z_conv1_fwd_gemm python tools/step_once.py atari 2 > $O/r2z_ncu_conv1.log 2>&1; echo "ncu rc=$?"; tail -2 $O/r2z_ncu_conv1.log
