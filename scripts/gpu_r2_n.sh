#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/exp_flags.py 1105228 base=0,prefetch=64 6 4 > gpurun_out/r2n_flags_1.1M.jsonl 2> gpurun_out/r2n.err; echo rc=$?; cat gpurun_out/r2n_flags_1.1M.jsonl
timeout 600 python scripts/exp_flags.py 8841823 base=0,prefetch=64 4 6 > gpurun_out/r2n_flags_c2.jsonl 2>> gpurun_out/r2n.err; echo rc=$?; cat gpurun_out/r2n_flags_c2.jsonl
tail -3 gpurun_out/r2n.err
