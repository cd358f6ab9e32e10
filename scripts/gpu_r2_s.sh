#!/bin/bash
# non-stationary corpus (--data shift: second half of the file drawn around another mean) under the sample-slab plans
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rows in 1105228 2210456 8841823; do
timeout 300 python bench.py --data shift --rows $rows --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2s_shift_$rows.json 2> gpurun_out/r2s_shift_$rows.err; echo "rows $rows rc=$?"
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/r2s_shift_$rows.json') if l.startswith('{')][-1])
print(j['ms_per_step'], j['engine'], j['parity_check'].get('ok'), j['parity_check'].get('hard_id_mismatches'))
PY
done
