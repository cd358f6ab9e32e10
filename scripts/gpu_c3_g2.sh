#!/bin/bash
# C3 shape (17 683 646 rows, the bilingual combined index) sharded over 2 GPUs
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --rows 17683646 --steps 5 --warmup 3 --stage-times > gpurun_out/bench_c3_g2.log 2>&1
echo "exit $?" >> gpurun_out/bench_c3_g2.log
grep '^{' gpurun_out/bench_c3_g2.log | cut -c1-300; grep stage_ms gpurun_out/bench_c3_g2.log | cut -c1-250
