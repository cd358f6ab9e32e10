#!/bin/bash
# full GPU suite + default bench with per-slab times
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
CMX_DEBUG_SLABS=1 timeout -s KILL 600 python bench.py --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench_iid.log 2> gpurun_out/bench_iid.err
echo "bench exit $?" >> gpurun_out/bench_iid.log
grep '^{' gpurun_out/bench_iid.log | cut -c1-300
tail -n 7 gpurun_out/bench_iid.err
