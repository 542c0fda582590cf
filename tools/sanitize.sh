#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (memcheck + racecheck at small sizes).
#   tools/sanitize.sh [outdir]      (run on a GPU box; logs -> outdir, default gpurun_out/)
# memcheck: the tiny commit/prove path, MSM edge cases, small NTTs; racecheck: the shared-memory kernels
# (NTT tiles, tree sums, scans); initcheck: uninitialised device reads of the MSM scratch.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool name pytest-args...
  local tool=$1 name=$2
  shift 2
  timeout 1500 $SAN --tool "$tool" --error-exitcode 86 --print-limit 20 \
      python -m pytest -x -q -m gpu -p no:cacheprovider "$@" > "$OUT/sanitizer_${tool}_${name}.log" 2>&1
  echo "exit=$?" >> "$OUT/sanitizer_${tool}_${name}.log"
  echo "== $tool $name"; tail -4 "$OUT/sanitizer_${tool}_${name}.log"
}
PROVE=(tests/test_gpu_prove.py -k "two_stage or batch_matches_single or single_stage")
MSM=(tests/test_gpu_msm.py -k "not full_size and not forked and not batch_affine")
NTT=(tests/test_gpu_ntt.py -k "matches_oracle or batch_and_device")
run memcheck prove "${PROVE[@]}"
run memcheck msm "${MSM[@]}"
run memcheck ntt "${NTT[@]}"
run racecheck prove "${PROVE[@]}"
run racecheck ntt "${NTT[@]}"
run initcheck prove "${PROVE[@]}"
