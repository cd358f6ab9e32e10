#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 900 python scripts/exp_paths.py > gpurun_out/exp_paths.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_paths.log
tail -n 4 gpurun_out/pytest_gpu.log; cat gpurun_out/exp_paths.log
