#!/bin/bash
# multi-GPU checks (run with gpurun --gpus N): correctness of the sharded paths, then the bench line at N
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
STEPS=${2:-10}
P=29500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P scripts/check_dist.py > gpurun_out/r2m_check_g$N.log 2>&1
echo "check_dist rc=$?"; grep -E "check_dist|Error|error" gpurun_out/r2m_check_g$N.log | tail -5
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "two_gpu or shards_index" > gpurun_out/r2m_tests_g$N.log 2>&1
echo "2gpu tests rc=$?"; tail -3 gpurun_out/r2m_tests_g$N.log
timeout 600 python scripts/exp_shards_one_process.py $N > gpurun_out/r2m_shards_one_process_g$N.json 2> gpurun_out/r2m_shards_one_process_g$N.err
echo "one-process rc=$?"; cat gpurun_out/r2m_shards_one_process_g$N.json; tail -3 gpurun_out/r2m_shards_one_process_g$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps $STEPS --warmup 3 --stage-times > gpurun_out/r2m_bench_g$N.json 2> gpurun_out/r2m_bench_g$N.err
echo "bench rc=$?"; grep -E "stage_ms|Error|error|Traceback" gpurun_out/r2m_bench_g$N.err | tail -5; head -c 2500 gpurun_out/r2m_bench_g$N.json
