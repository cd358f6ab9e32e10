#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 120 -x -k "single_and_pair" > gpurun_out/pytest_pair.log 2>&1
echo "pytest pair exit $?" >> gpurun_out/pytest_pair.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 900 python scripts/exp_tc.py 1000000,4420912,8841823 > gpurun_out/exp_tc.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_tc.log
tail -n 4 gpurun_out/pytest_pair.log gpurun_out/pytest_gpu.log; cat gpurun_out/exp_tc.log
