#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# corpora that end after the dense slab / after the mid slab / the whole shard: differences attribute the cycles
CMX_LIB=$PWD/codemix-dense-retrieval_b200/lib/libcmx_timers.so timeout 300 python scripts/exp_tc_timers.py 2048 85760 1105228 > gpurun_out/r2m_timers.jsonl 2> gpurun_out/r2m_timers.err; echo rc=$?
cat gpurun_out/r2m_timers.jsonl; tail -3 gpurun_out/r2m_timers.err
