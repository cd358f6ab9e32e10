#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/exp_sched.py 8841823 > gpurun_out/r2f_sched_c2.jsonl 2>gpurun_out/r2f_sched.err; cat gpurun_out/r2f_sched_c2.jsonl; tail -3 gpurun_out/r2f_sched.err
timeout 600 python scripts/exp_sched.py 1105228 > gpurun_out/r2f_sched_1.1M.jsonl 2>>gpurun_out/r2f_sched.err; cat gpurun_out/r2f_sched_1.1M.jsonl
timeout 600 python scripts/ab_lib.py 1105228 6980 1000 > gpurun_out/r2f_ab_shard1.1M.txt 2>&1; grep RESULT gpurun_out/r2f_ab_shard1.1M.txt
