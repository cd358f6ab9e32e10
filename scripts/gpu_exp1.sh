#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 900 python scripts/exp_paths.py > gpurun_out/exp_paths.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_paths.log
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_full.log
tail -n 3 gpurun_out/pytest_gpu.log
