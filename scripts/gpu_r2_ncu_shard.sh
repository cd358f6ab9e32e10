#!/bin/bash
# ncu --set full of the scoring launches of an 8-GPU-shard-sized search (1 105 228 rows) on one GPU: short launches, not power-capped
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --rows 1105228 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$CMD > gpurun_out/r2s_plain.log 2>&1 && grep '^{' gpurun_out/r2s_plain.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 11 -c 1 -f -o gpurun_out/prof_tc_shard_r2 $CMD > gpurun_out/r2s_ncu.log 2>&1
echo "capture exit $?"
CMX_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:tc_score_pair -s 3 -c 1 -f -o gpurun_out/prof_tcpair_shard_r2 $CMD > gpurun_out/r2s_ncu_pair.log 2>&1
echo "pair capture exit $?"
