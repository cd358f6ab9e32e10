#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "tensor_search_parity or rescore or anisotropic" > gpurun_out/pytest_quick.log 2>&1
echo "pytest quick exit $?" >> gpurun_out/pytest_quick.log
timeout -s KILL 900 python scripts/exp_tc.py 1000000,8841823 > gpurun_out/exp_tc.log 2>&1
echo "exp exit $?" >> gpurun_out/exp_tc.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_quick.log; cat gpurun_out/exp_tc.log; tail -n 15 gpurun_out/pytest_gpu.log
