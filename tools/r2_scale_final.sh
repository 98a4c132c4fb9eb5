# Round 2 scaling check on one 8-GPU box: N = 8 and N = 4 through the driver's own launch line, the 2-rank NCCL test.
O=gpurun_out
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > $O/r3d_bench_$n.json 2> $O/r3d_bench_$n.err; echo "N=$n rc=$? $(grep -h '^{' $O/r3d_bench_$n.json | cut -c1-220)"; tail -2 $O/r3d_bench_$n.err
done
timeout 300 python -m pytest tests/test_parallel_gpu.py -q > $O/r3d_nccl_test.log 2>&1; echo "nccl test rc=$?"; tail -2 $O/r3d_nccl_test.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/r3d_bench_1.json 2> $O/r3d_bench_1.err; echo "N=1 rc=$? $(cut -c1-160 $O/r3d_bench_1.json)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r3d_bench_2.json 2> $O/r3d_bench_2.err; echo "N=2 rc=$? $(grep -h '^{' $O/r3d_bench_2.json | cut -c1-160)"
