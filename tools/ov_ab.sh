# A/B harness used while tuning: bench.py under a few environment settings, one JSON per setting in gpurun_out/.
# usage (on the GPU box): bash tools/ov_ab.sh "PD_B200_OVERLAP=0" "PD_B200_OVERLAP=3" ...
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 200 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --watchdog 120 \
      > gpurun_out/ab_bench_$i.json 2> gpurun_out/ab_bench_$i.err
  echo "CFG $i [$cfg] rc=$? $(cut -c1-120 gpurun_out/ab_bench_$i.json)" | tee -a gpurun_out/ab_summary.txt
done
