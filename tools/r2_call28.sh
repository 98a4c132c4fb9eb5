O=gpurun_out
S="2402500,48,48,0,0,0,1 2250000,48,108,0,1 2250000,108,48 48,48,2402500,1,1,1 108,48,2250000,1,1,1"
timeout 300 python tools/gemm_bench.py $S --reps 5 > $O/r2z_small_nk.jsonl 2>&1; cat $O/r2z_small_nk.jsonl | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pd_gemm --launch-skip 15 --launch-count 5 -f -o $O/r2z_small_nk python tools/gemm_bench.py $S --reps 1 > $O/r2z_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/r2z_ncu.log
