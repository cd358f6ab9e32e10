#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 900 python scripts/exp_tc.py 8841823 > gpurun_out/exp_tc.log 2>&1
echo "exit $?" >> gpurun_out/exp_tc.log
grep -v Warning gpurun_out/exp_tc.log | tail -n 12
