#!/bin/bash
# coalesced dense epilogue: parity tests, A/B against the previous build (shard, C2, and k=100 where the dense slab is 8192 rows)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2r_tests.log
L=codemix-dense-retrieval_b200/lib
timeout 300 python scripts/ab_lib.py 1105228 6980 1000 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2r_ab_shard1.1M.txt 2>&1; grep -v "^ *$" gpurun_out/r2r_ab_shard1.1M.txt | grep -v "slab [12]"
timeout 300 python scripts/ab_lib.py 1105228 6980 100 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2r_ab_shard1.1M_k100.txt 2>&1; grep -v "^ *$" gpurun_out/r2r_ab_shard1.1M_k100.txt | grep -v "slab [123]"
timeout 600 python scripts/ab_lib.py 8841823 6980 1000 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2r_ab_c2.txt 2>&1; grep -v "^ *$" gpurun_out/r2r_ab_c2.txt | grep -v "slab [12]"
