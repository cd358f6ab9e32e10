#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_round2.py -m gpu -q --timeout 300 -x > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -c 600 gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2d_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['ms_per_step'],'frac',j['roofline']['frac'],'kernel_ms',j['roofline']['kernel_ms_per_step'],'select',j['select_ms_per_step'])
for e in j['extra_configs']['C4_small_batch']:
    print(e['nq'],e['k'],round(e['ms_per_search'],3),round(e['roofline']['kernel_ms_per_step'],3),round(e['roofline']['frac'],3))
print({k:(round(v['ms_per_step'],2) if 'ms_per_step' in v else round(v['ms_per_job'],1)) for k,v in j['extra_configs'].items() if isinstance(v,dict)})
PY
