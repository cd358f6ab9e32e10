#!/bin/bash
# multi-GPU quick: correctness check + default bench at N = $1
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 scripts/check_dist.py > gpurun_out/check_dist_g$N.log 2>&1
echo "check_dist exit $?" >> gpurun_out/check_dist_g$N.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --stage-times > gpurun_out/bench_g$N.log 2>&1
echo "bench g$N exit $?" >> gpurun_out/bench_g$N.log
grep check_dist gpurun_out/check_dist_g$N.log | cut -c1-300; grep '^{' gpurun_out/bench_g$N.log | cut -c1-1500
