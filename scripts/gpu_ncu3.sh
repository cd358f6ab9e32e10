#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
CMD1="python scripts/exp_tc.py 2000000"
$CMD1 > gpurun_out/plain_exp_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_score -s 60 -c 12 -f -o gpurun_out/prof_tc3 $CMD1 > gpurun_out/ncu_tc3.log 2>&1
echo "tc capture exit $?" >> gpurun_out/ncu_tc3.log
cat gpurun_out/plain_exp_tc.log
