#!/bin/bash
# re-run GPU tests, then ncu launch list + full capture of the dominant kernels (1 GPU)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
KRE='regex:tc_score|stream_score|compact_kernel|mix_normalize|split_planes|absmax|ws_init|set_counts|merge_kernel|scale_from'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches_full.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD1 > gpurun_out/plain_full1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_score_kernel -s 9 -c 3 -f -o gpurun_out/prof_tc $CMD1 > gpurun_out/ncu_tc.log 2>&1
echo "tc capture exit $?" >> gpurun_out/ncu_tc.log
CMD2="python bench.py --nq 8 --k 100 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_stream.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stream_score -s 4 -c 2 -f -o gpurun_out/prof_stream $CMD2 > gpurun_out/ncu_stream.log 2>&1
echo "stream capture exit $?" >> gpurun_out/ncu_stream.log
ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 6 -c 2 -f -o gpurun_out/prof_compact $CMD1 > gpurun_out/ncu_compact.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log
