#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=codemix-dense-retrieval_b200/lib
CMX_LIB=$PWD/$L/libcmx_timers.so timeout 300 python scripts/exp_tc_timers.py 1105228 > gpurun_out/r2o_timers.jsonl 2> gpurun_out/r2o_timers.err; echo rc=$?
cat gpurun_out/r2o_timers.jsonl; tail -3 gpurun_out/r2o_timers.err
timeout 600 python scripts/ab_lib.py 1105228 6980 1000 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2o_ab_shard1.1M.txt 2>&1; grep -v "^ *$" gpurun_out/r2o_ab_shard1.1M.txt
timeout 600 python scripts/ab_lib.py 8841823 6980 1000 $L/libcmx_prev.so $L/libcmx.so > gpurun_out/r2o_ab_c2.txt 2>&1; grep -v "^ *$" gpurun_out/r2o_ab_c2.txt
