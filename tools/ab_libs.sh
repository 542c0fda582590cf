#!/bin/bash
# A/B the library variants built with csrc/Makefile OUT=/BUILD=/EXTRA= (same ABI; bench.py picks HK_LIB).
# usage: tools/ab_libs.sh tag lib1.so lib2.so ...   -> gpurun_out/ab_<tag>_<lib>.json
tag=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  HK_LIB=$PWD/hekaton_system_b200/lib/$lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline \
      > gpurun_out/ab_${tag}_${name}.json 2> gpurun_out/ab_${tag}_${name}.err
  python - "$name" gpurun_out/ab_${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    x = d.get("extras", {})
    print(sys.argv[1], "proofs/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "accum frac", round(d["roofline"]["frac"], 3),
          "accum ms", round(d["roofline"]["avg_launch_ms"], 3), "msm2^20 ms", round(x.get("msm_g1_2p20", {}).get("ms", 0), 3),
          "verified", d.get("verified_vs_trapdoor"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
