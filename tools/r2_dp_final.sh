# closing data-parallel check on two GPUs: the 2-rank NCCL test and the driver-style N = 2 bench line
O=gpurun_out
timeout 300 python -m pytest tests/test_parallel_gpu.py -q > $O/r3z_nccl_test.log 2>&1; echo "nccl test rc=$?"; tail -2 $O/r3z_nccl_test.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r3z_bench_2.json 2> $O/r3z_bench_2.err; echo "N=2 rc=$? $(grep -h '^{' $O/r3z_bench_2.json | cut -c1-260)"
grep -h '^{' $O/r3z_bench_2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['clocks'], d.get('gpu_launches'))"
