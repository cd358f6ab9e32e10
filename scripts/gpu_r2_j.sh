#!/bin/bash
# value-binned selection (block_kth_largest fast path): parity tests, then A/B against the previous build on the same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2j_tests.log
L=codemix-dense-retrieval_b200/lib
timeout 600 python scripts/ab_lib.py 1105228 6980 1000 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2j_ab_shard1.1M.txt 2>&1; grep -v "^ *$" gpurun_out/r2j_ab_shard1.1M.txt
timeout 600 python scripts/ab_lib.py 8841823 6980 1000 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2j_ab_c2.txt 2>&1; grep RESULT gpurun_out/r2j_ab_c2.txt
