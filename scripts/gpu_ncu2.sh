#!/bin/bash
# ncu evidence for the final kernels (1 GPU): launch list of the default bench + full captures
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
KRE='regex:tc_score|stream_score|compact_kernel|mix_normalize|split_planes|absmax|ws_init|set_counts|merge|scale_from'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches_full.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
# full capture of the six scoring launches of one step on a half-size shard (keeps every launch < 100 ms:
# the hardware counters of `--set full` overflowed on the 150 ms launch of the full-size run)
CMD1="python bench.py --rows 4420912 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD1 > gpurun_out/plain_half.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 6 -c 6 -f -o gpurun_out/prof_tc2 $CMD1 > gpurun_out/ncu_tc2.log 2>&1
echo "tc capture exit $?" >> gpurun_out/ncu_tc2.log
CMD2="python bench.py --nq 4 --k 100 --path stream --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_stream.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stream_score -s 4 -c 3 -f -o gpurun_out/prof_stream2 $CMD2 > gpurun_out/ncu_stream2.log 2>&1
echo "stream capture exit $?" >> gpurun_out/ncu_stream2.log
ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 6 -c 6 -f -o gpurun_out/prof_compact2 $CMD1 > gpurun_out/ncu_compact2.log 2>&1
echo "compact capture exit $?" >> gpurun_out/ncu_compact2.log
# the default bench, plain, for the record
python bench.py > gpurun_out/bench_default.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_default.log
python bench.py --nq 4 --k 100 --path stream --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stream_nq4.log 2>&1
python bench.py --nq 1 --k 100 --path stream --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stream_nq1.log 2>&1
tail -n 2 gpurun_out/bench_default.log | cut -c1-400
