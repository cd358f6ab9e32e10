#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_configs.py -m gpu -q --timeout 300 -x > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?"; tail -c 800 gpurun_out/r2c_tests.log
timeout 900 python scripts/exp_prescore.py 0,1.5,0,8 2,1.5,0,8 1,1.5,0,8 1,1.0,12288,8 1,1.0,0,4 1,0.7,0,8 0,1.5,0,8 1,1.0,0,8 > gpurun_out/r2c_prescore.jsonl 2> gpurun_out/r2c_prescore.err
echo "prescore rc=$?"; cat gpurun_out/r2c_prescore.jsonl; tail -c 600 gpurun_out/r2c_prescore.err
