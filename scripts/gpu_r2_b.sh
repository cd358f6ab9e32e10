#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py "tests/test_gpu_parity.py::test_merge_and_sharded_equal_single" tests/test_gpu_cli.py -m gpu -q --timeout 300 > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -c 1500 gpurun_out/r2b_tests.log
timeout 900 python scripts/exp_prescore.py > gpurun_out/r2b_prescore.jsonl 2> gpurun_out/r2b_prescore.err
echo "prescore rc=$?"; cat gpurun_out/r2b_prescore.jsonl; tail -c 600 gpurun_out/r2b_prescore.err
timeout 600 python scripts/exp_loader.py > gpurun_out/r2b_loader.jsonl 2> gpurun_out/r2b_loader.err
echo "loader rc=$?"; cat gpurun_out/r2b_loader.jsonl; tail -c 600 gpurun_out/r2b_loader.err
df -h /tmp /dev/shm | tee gpurun_out/r2b_df.txt
