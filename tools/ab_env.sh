#!/bin/bash
# A/B environment-variable knobs of one library build.  usage: tools/ab_env.sh tag "VAR=val ..." "VAR2=val" ...
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/abenv_${tag}_$i.json 2> gpurun_out/abenv_${tag}_$i.err
  python - "$envs" gpurun_out/abenv_${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    x = d.get("extras", {})
    print(sys.argv[1], "| proofs/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "accum frac", round(d["roofline"]["frac"], 3),
          "accum ms", round(d["roofline"]["avg_launch_ms"], 3), "msm2^20 ms", round(x.get("msm_g1_2p20", {}).get("ms", 0), 3),
          "verified", d.get("verified_vs_trapdoor"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
