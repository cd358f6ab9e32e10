#!/bin/bash
# round-2 profiler evidence (ONE GPU): launch list of the default bench command, full captures of the dominant
# kernel (the three scoring launches of one step), the rescoring kernel and the small-batch scorer.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
KRE='regex:tc_score|stream_score|compact_kernel|rescore_kernel|prescore|mix_normalize|split_planes|absmax|ws_init|set_counts|merge|union_kth|export_scores|scale_from|row_norm|row_resid|query_margin|snapshot|publish|store2|max_bounds'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$CMD > gpurun_out/r2n_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/launches_full.csv $CMD > gpurun_out/r2n_ncu_launches.log 2>&1
echo "launch list exit $?"
grep '^{' gpurun_out/r2n_plain.log | cut -c1-300
# full captures: launches 3.. (after the warm-up searches) = dense, mid, final slab of one step
ncu --set full --clock-control none --import-source on -k regex:tc_score_ -s 9 -c 3 -f -o gpurun_out/prof_tc_r2 $CMD > gpurun_out/r2n_ncu_tc.log 2>&1
echo "tc capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:rescore_kernel -s 3 -c 1 -f -o gpurun_out/prof_rescore_r2 $CMD > gpurun_out/r2n_ncu_rescore.log 2>&1
echo "rescore capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 9 -c 3 -f -o gpurun_out/prof_compact_r2 $CMD > gpurun_out/r2n_ncu_compact.log 2>&1
echo "compact capture exit $?"
CMDS="python bench.py --nq 16 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$CMDS > gpurun_out/r2n_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_score_small -s 9 -c 3 -f -o gpurun_out/prof_small_r2 $CMDS > gpurun_out/r2n_ncu_small.log 2>&1
echo "small capture exit $?"
ls -la gpurun_out/*.ncu-rep
