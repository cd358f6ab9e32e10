#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/exp_variants.py 1105228 single256,pair 6 4 > gpurun_out/r2p_variants_1.1M.jsonl 2> gpurun_out/r2p.err; echo rc=$?; cat gpurun_out/r2p_variants_1.1M.jsonl
timeout 300 python scripts/exp_variants.py 2210456 single256,pair 5 4 > gpurun_out/r2p_variants_2.2M.jsonl 2>> gpurun_out/r2p.err; echo rc=$?; cat gpurun_out/r2p_variants_2.2M.jsonl
tail -2 gpurun_out/r2p.err
