#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python scripts/ab_lib.py 1105228 6980 1000 > gpurun_out/r2e_ab_shard1.1M.txt 2>&1; cat gpurun_out/r2e_ab_shard1.1M.txt
timeout 900 python scripts/ab_lib.py 8841823 6980 1000 > gpurun_out/r2e_ab_c2.txt 2>&1; cat gpurun_out/r2e_ab_c2.txt
