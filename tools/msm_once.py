#!/usr/bin/env python3
"""Run a few MSMs (window-table mode) -- target for ncu captures.
usage: msm_once.py LOG_N REPS GROUP KIND(uniform|witness) K DENSITY      (HK_PROBE_OPTS="g2pair=1,smin=32": hk_ctx_set_option knobs)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from hekaton_system_b200 import codec
from hekaton_system_b200.backend import Context

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
group = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kind = sys.argv[4] if len(sys.argv) > 4 else "uniform"
K = int(sys.argv[5]) if len(sys.argv) > 5 else 1
density = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = Context("bls12_381", 0, stream=stream.cuda_stream)
for kv in filter(None, os.environ.get("HK_PROBE_OPTS", "").split(",")):
    from msm_time_names import NAMES
    ctx.set_option(NAMES[kv.split("=")[0]], int(kv.split("=")[1]))
cv = ctx.curve
n = 1 << lg
rng = np.random.default_rng(1)
ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
ks[:, 3] >>= 2
gen = codec.g1_affine_to_limbs([cv.g1], cv)[0] if group == 1 else codec.g2_affine_to_limbs([cv.g2], cv)[0]
h = ctx.bases_from_scalars(group, gen, ks)
if os.environ.get("HK_PROBE_TABLE", "1") == "1":
    ctx.bases_precompute(h, density)
sc = rng.integers(0, 1 << 63, size=(K * n, 4), dtype=np.uint64)
sc[:, 3] >>= 2
if kind == "witness":          # 10 % full-width, 45 % ones, 45 % zeros (the bench workload's mix)
    u = rng.random(K * n)
    sc[u >= 0.1] = 0
    sc[(u >= 0.1) & (u < 0.55), 0] = 1
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
d_out = torch.zeros(64 * K, dtype=torch.int64, device="cuda")
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.msm(group, h, d_sc.data_ptr(), n=n, K=K, out=d_out.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"msm g{group} 2^{lg} K={K} {kind}: {e0.elapsed_time(e1):.3f} ms")
ctx.close()
