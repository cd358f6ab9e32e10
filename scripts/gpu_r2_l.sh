#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "parity tests rc=$?"; tail -3 gpurun_out/r2l_tests.log
CMX_LIB=$PWD/codemix-dense-retrieval_b200/lib/libcmx_timers.so timeout 300 python scripts/exp_tc_timers.py 2048 85760 1105228 > gpurun_out/r2m_timers.jsonl 2> gpurun_out/r2m_timers.err; echo rc=$?
cat gpurun_out/r2m_timers.jsonl; tail -3 gpurun_out/r2m_timers.err
timeout 300 python scripts/exp_epilogue.py 1105228 > gpurun_out/r2k_epi.jsonl 2> gpurun_out/r2k_epi.err; echo rc=$?
cat gpurun_out/r2k_epi.jsonl
grep -A3 "round 3" gpurun_out/r2k_epi.err | grep -v "^--" | head -40
