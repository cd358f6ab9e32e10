#!/bin/bash
# ncu launch list of the default bench command (after it has run plain)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
KRE='regex:tc_score|stream_score|compact_kernel|rescore_kernel|mix_normalize|split_planes|absmax|ws_init|set_counts|merge|scale_from|row_norm|row_resid|query_margin'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/launches_full.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
grep '^{' gpurun_out/plain_full.log | cut -c1-300
tail -n 2 gpurun_out/ncu_launches.log
