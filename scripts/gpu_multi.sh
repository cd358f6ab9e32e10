#!/bin/bash
# multi-GPU: tests that need >=2 GPUs, then torchrun bench at N = $1
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -k "two_gpu or sharded or merge" > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 scripts/check_dist.py > gpurun_out/check_dist_g$N.log 2>&1
echo "check_dist exit $?" >> gpurun_out/check_dist_g$N.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --exchange allgather > gpurun_out/bench_g${N}_allgather.log 2>&1
echo "bench allgather g$N exit $?" >> gpurun_out/bench_g${N}_allgather.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_g$N.log 2>&1
echo "bench g$N exit $?" >> gpurun_out/bench_g$N.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_g$N.log 2>&1
echo "bench ref g$N exit $?" >> gpurun_out/bench_ref_g$N.log
tail -n 3 gpurun_out/pytest_multi.log; grep check_dist gpurun_out/check_dist_g$N.log
