#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 900 python scripts/exp_sweep.py > gpurun_out/exp_sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/exp_sweep.log
grep -v Warning gpurun_out/exp_sweep.log | tail -n 5
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -k "cli or sweep or dropin" > gpurun_out/pytest_cli.log 2>&1
tail -n 3 gpurun_out/pytest_cli.log
