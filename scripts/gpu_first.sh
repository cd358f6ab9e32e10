#!/bin/bash
# first GPU contact: non-tensor tests, then tensor tests, then a reduced and the full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -k "not tensor and not large and not sharded and not mixed and not sweep and not dropin and not adversarial and not incremental and not device_tensors and not ragged" > gpurun_out/pytest_stream.log 2>&1
echo "stream-ish tests exit $?" >> gpurun_out/pytest_stream.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -k "tensor or large or sharded or mixed or sweep or dropin or adversarial or incremental or device_tensors or ragged" > gpurun_out/pytest_tensor.log 2>&1
echo "tensor tests exit $?" >> gpurun_out/pytest_tensor.log
timeout -s KILL 600 python bench.py --rows 1000000 --steps 3 --warmup 3 > gpurun_out/bench_1m.log 2>&1
echo "bench 1m exit $?" >> gpurun_out/bench_1m.log
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench full exit $?" >> gpurun_out/bench_full.log
timeout -s KILL 300 python bench.py --rows 8841823 --nq 8 --k 100 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stream.log 2>&1
echo "bench stream exit $?" >> gpurun_out/bench_stream.log
tail -5 gpurun_out/pytest_stream.log gpurun_out/pytest_tensor.log gpurun_out/bench_1m.log gpurun_out/bench_full.log gpurun_out/bench_stream.log
