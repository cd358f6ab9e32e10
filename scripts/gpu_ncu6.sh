#!/bin/bash
# final round-1 evidence: full GPU suite, plain bench
# lines, launch list, full captures
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for nq in 1 16 32; do
  python bench.py --nq $nq --k 100 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nq$nq.log 2>&1
  grep '^{' gpurun_out/bench_nq$nq.log | cut -c1-200
done
python bench.py > gpurun_out/bench_default.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_default.log
python bench.py --precision split --no-cpu-baseline > gpurun_out/bench_split.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
KRE='regex:tc_score|stream_score|compact_kernel|rescore_kernel|mix_normalize|split_planes|absmax|ws_init|set_counts|merge|scale_from|row_norm|row_resid|query_margin'
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/launches_full.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
CMD1="python bench.py --rows 4420912 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD1 > gpurun_out/plain_half.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_score -s 7 -c 7 -f -o gpurun_out/prof_tc6 $CMD1 > gpurun_out/ncu_tc6.log 2>&1
echo "tc capture exit $?" >> gpurun_out/ncu_tc6.log
ncu --set full --clock-control none --import-source on -k regex:rescore_kernel -s 1 -c 1 -f -o gpurun_out/prof_rescore6 $CMD1 > gpurun_out/ncu_rescore6.log 2>&1
echo "rescore capture exit $?" >> gpurun_out/ncu_rescore6.log
grep '^{' gpurun_out/bench_default.log | cut -c1-1800
grep '^{' gpurun_out/bench_split.log | cut -c1-300
grep '^{' gpurun_out/bench_reference.log | cut -c1-600
