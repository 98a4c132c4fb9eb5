O=gpurun_out
timeout 300 python -m pytest tests/test_parallel_gpu.py -q > $O/r2t_nccl_test.log 2>&1; echo "nccl test rc=$?"; tail -3 $O/r2t_nccl_test.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2t_bench_2.json 2> $O/r2t_bench_2.err; echo "N=2 rc=$? $(grep -h '^{' $O/r2t_bench_2.json | cut -c1-160)"
