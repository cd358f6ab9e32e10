#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?"; tail -c 700 gpurun_out/r2g_tests.log
python scripts/diag_chunks_1gpu.py 8 5,6,7 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2g_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['ms_per_step'],'frac',j['roofline']['frac'],'kernel_ms',j['roofline']['kernel_ms_per_step'],'select',j['select_ms_per_step'], j['memory_gb_per_gpu'], j['parity_check']['ok'], j['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_reference.json 2> gpurun_out/r2g_reference.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/r2g_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
