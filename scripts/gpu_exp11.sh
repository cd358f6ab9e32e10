#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python scripts/exp_slabs.py 1105228 > gpurun_out/exp_slabs.log 2>&1
tail -n 70 gpurun_out/exp_slabs.log
