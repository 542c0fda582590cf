#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (memcheck + racecheck at small sizes).
#   tools/sanitize.sh [outdir]      (run on a GPU box; logs -> outdir, default gpurun_out/)
# memcheck: the tiny commit/prove path, MSM edge cases, small NTTs; racecheck: the shared-memory kernels
# (NTT tiles, tree sums, scans).  initcheck: uninitialised device reads of the MSM scratch.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL_PROVE='tests/test_gpu_prove.py::test_two_stage_commit_prove tests/test_gpu_prove.py::test_prove_batch_matches_single tests/test_gpu_prove.py::test_single_stage_and_zero_randomness'
SEL_MSM='tests/test_gpu_msm.py -k "not full_size and not forked and not batch_affine"'
SEL_NTT='tests/test_gpu_ntt.py -k "matches_oracle or batch_and_device"'
run() {  # tool, name, selection
  timeout 1500 $SAN --tool "$1" --error-exitcode 86 --print-limit 20 \
      python -m pytest -x -q -m gpu -p no:cacheprovider $3 > "$OUT/sanitizer_$1_$2.log" 2>&1
  echo "exit=$?" >> "$OUT/sanitizer_$1_$2.log"
  tail -4 "$OUT/sanitizer_$1_$2.log"
}
eval run memcheck prove "'$SEL_PROVE'"
eval run memcheck msm "'$SEL_MSM'"
eval run memcheck ntt "'$SEL_NTT'"
eval run racecheck prove "'$SEL_PROVE'"
eval run racecheck ntt "'$SEL_NTT'"
eval run initcheck prove "'$SEL_PROVE'"
