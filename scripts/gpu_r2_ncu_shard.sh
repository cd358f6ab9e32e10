#!/bin/bash
# ncu --set full on an 8-GPU-shard-sized search (1 105 228 rows) on one GPU: the scoring launches of one step
# (dense first slab, speculative mid slab, final slab) and the three compactions
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --rows 1105228 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$CMD > gpurun_out/r2s_plain.log 2>&1 && grep '^{' gpurun_out/r2s_plain.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 9 -c ${1:-2} -f -o gpurun_out/prof_tc_shard_r2 $CMD > gpurun_out/r2s_ncu.log 2>&1
echo "score capture exit $?"
if [ "${2:-0}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 9 -c 3 -f -o gpurun_out/prof_compact_shard_r2 $CMD > gpurun_out/r2s_ncu_c.log 2>&1
echo "compact capture exit $?"
fi
