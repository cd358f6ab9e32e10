#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 900 python scripts/exp_sweep.py > gpurun_out/exp_sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/exp_sweep.log
timeout -s KILL 900 python bench.py --rows 17683646 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_1gpu.log 2>&1
echo "c3 exit $?" >> gpurun_out/bench_c3_1gpu.log
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/exp_sweep.log; tail -n 2 gpurun_out/bench_c3_1gpu.log | cut -c1-700
