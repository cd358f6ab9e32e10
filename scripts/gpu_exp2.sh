#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1200 python scripts/exp_tc.py > gpurun_out/exp_tc.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_tc.log
tail -n 30 gpurun_out/exp_tc.log
