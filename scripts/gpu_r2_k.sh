#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/exp_epilogue.py 1105228 > gpurun_out/r2k_epi.jsonl 2> gpurun_out/r2k_epi.err; echo rc=$?
cat gpurun_out/r2k_epi.jsonl
grep -A4 "round 3" gpurun_out/r2k_epi.err | grep -v "^--" | head -40
