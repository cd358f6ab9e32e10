#!/bin/bash
# per-call cost at the reference's qblock sizes (onepass_dense_mix_run_custom_lang.py:424, run_all_vector_pairs.sh:779) and k
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for cfg in "128 100" "256 100" "1024 100" "6980 100" "6980 500"; do
  set -- $cfg
  python bench.py --nq $1 --k $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nq$1_k$2.log 2>&1
  grep '^{' gpurun_out/bench_nq$1_k$2.log | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print(j['config']['queries'], j['config']['k'], round(j['value']), 'q/s', round(j['ms_per_step'], 3), 'ms', j['roofline']['bound'], round(j['roofline']['frac'], 3), 'e2e', round(j['e2e']['value']))
"
done
