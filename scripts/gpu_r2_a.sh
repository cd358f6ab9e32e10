#!/bin/bash
# round 2, first contact: new tests, then the whole GPU suite, then a short bench (1 GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
free -g > gpurun_out/r2a_mem.txt 2>&1; nproc >> gpurun_out/r2a_mem.txt
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_eps_bound.py -m gpu -q -s --timeout 300 > gpurun_out/r2a_new_tests.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/r2a_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_round2.py --deselect tests/test_gpu_eps_bound.py > gpurun_out/r2a_all_tests.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/r2a_all_tests.log
CMX_DEBUG_SLABS=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_new_tests.log; tail -c 400 gpurun_out/r2a_all_tests.log; tail -c 1500 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
