#!/bin/bash
# compute-sanitizer passes over the small-shape parity tests: memcheck, racecheck (shared-memory
# hazards in the hand-rolled mbarrier / TMEM pipelines: single-CTA, CTA-pair and small-batch scoring kernels, the
# radix-select kernels) and synccheck.  Run under gpurun on ONE GPU:
#     gpurun --timeout 2400 -- 'bash scripts/gpu_sanitize.sh'
# Logs land in gpurun_out/sanitize_<tool>.log; the summary line of each tool is echoed at the end.
# NOTE (round 2): this pool answers `compute-sanitizer is closed on this pool and stays closed` (rc 86,
# profiles/r02_sanitizer_closed_on_pool.txt); the substitute that does run is
# tests/test_gpu_round2.py::test_repeated_runs_are_bit_identical_across_kernel_variants.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
# small shapes only: the sanitizer slows kernels by 10-100x
TESTS=(
  "tests/test_gpu_parity.py::test_tensor_single_and_pair_kernels"
  "tests/test_gpu_parity.py::test_tensor_small_batch_kernel"
  "tests/test_gpu_parity.py::test_stream_search_parity"
  "tests/test_gpu_parity.py::test_duplicate_rows_tie_order"
  "tests/test_gpu_parity.py::test_union_kth_matches_numpy"
  "tests/test_gpu_parity.py::test_two_phase_sharded_search_on_one_gpu"
  "tests/test_gpu_parity.py::test_mix_golden_bit_exact"
  "tests/test_gpu_round2.py::test_pending_two_phase_state_is_cancelled_by_other_calls"
  "tests/test_gpu_round2.py::test_empty_and_unusable_shards_in_two_phase"
)
SEL='(5003 or 9000 or nq5 or 17- or G2 or golden or duplicate or union or pending or empty or (stream and 3-10)) and not 70001'
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitize_${tool}.log
  timeout 1500 $SAN --tool $tool --target-processes all --error-exitcode 99 --print-limit 20 \
      python -m pytest "${TESTS[@]}" -m gpu -q -x --timeout 1200 -k "$SEL" > "$log" 2>&1
  echo "$tool rc=$?" | tee -a "$log"
done
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitize_${tool}.log | tail -4
done
