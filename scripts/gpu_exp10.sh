#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 900 python scripts/exp_tc.py 8841823 > gpurun_out/exp_tc.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_tc.log
timeout -s KILL 900 python scripts/exp_paths.py > gpurun_out/exp_paths.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_paths.log
cat gpurun_out/exp_tc.log gpurun_out/exp_paths.log
