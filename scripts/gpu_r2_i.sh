#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q --timeout 400 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 300 gpurun_out/r2i_bench.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print("value",j["value"],"ms",j["ms_per_step"],"e2e",j["e2e"]["ms_per_step"],"frac",j["roofline"]["frac"],"kernel_ms",j["roofline"]["kernel_ms_per_step"],"select",j["select_ms_per_step"], j["parity_check"]["ok"], j["parity_check"]["id_exact_frac"])
print({k:(round(v["ms_per_step"],2) if "ms_per_step" in v else (round(v["ms_per_job"],1) if "ms_per_job" in v else None)) for k,v in j["extra_configs"].items() if isinstance(v,dict)})
PY
