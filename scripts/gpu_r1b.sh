#!/bin/bash
# block-order + measured-margin build: full GPU suite, default bench, data variants, C3 shape
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
for v in iid aniso shift; do
  extra="--no-cpu-baseline"; [ $v = iid ] && extra=""
  CMX_DEBUG_SLABS=1 timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --data $v $extra > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err
  echo "bench $v exit $?" >> gpurun_out/bench_$v.log
  grep '^{' gpurun_out/bench_$v.log | cut -c1-400
  grep -o '"reruns": [0-9]*' gpurun_out/bench_$v.log
done
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 --rows 17683646 --no-cpu-baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err
echo "bench c3 exit $?" >> gpurun_out/bench_c3.log
grep '^{' gpurun_out/bench_c3.log | cut -c1-400
