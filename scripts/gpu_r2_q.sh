#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMX_LIB=$PWD/codemix-dense-retrieval_b200/lib/libcmx_timers.so timeout 300 python scripts/exp_pool_paths.py > gpurun_out/r2q_pool_paths.jsonl 2> gpurun_out/r2q.err; echo rc=$?; cat gpurun_out/r2q_pool_paths.jsonl; tail -3 gpurun_out/r2q.err
