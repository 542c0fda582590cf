#!/usr/bin/env python3
"""Kernel timeline (torch.profiler / CUPTI) of one standalone operation -- diagnostic.
usage: op_timeline.py msm LOG_N [GROUP] | ntt LOG_N"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from hekaton_system_b200 import codec
from hekaton_system_b200.backend import Context

op, lg = sys.argv[1], int(sys.argv[2])
group = int(sys.argv[3]) if len(sys.argv) > 3 else 1
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = Context("bls12_381", 0, stream=stream.cuda_stream)
cv = ctx.curve
rng = np.random.default_rng(1)
n = 1 << lg
if op == "msm":
    ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    ks[:, 3] >>= 2
    gen = codec.g1_affine_to_limbs([cv.g1], cv)[0] if group == 1 else codec.g2_affine_to_limbs([cv.g2], cv)[0]
    h = ctx.bases_from_scalars(group, gen, ks)
    ctx.bases_precompute(h, 1.0)
    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    sc[:, 3] >>= 2
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(128, dtype=torch.int64, device="cuda")
    run = lambda: ctx.msm(group, h, d_sc.data_ptr(), n=n, out=d_out.data_ptr())
else:
    a = rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    run = lambda: ctx.ntt(d.data_ptr(), lg)
for _ in range(2):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as tp:
    run()
    torch.cuda.synchronize()
tmp = "/tmp/_tl.json"
tp.export_chrome_trace(tmp)
evs = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
t0 = min(e["ts"] for e in evs)
for e in sorted(evs, key=lambda e: e["ts"]):
    print(f"{e['ts'] - t0:10.1f} {e['dur']:9.1f} {e['name'][:90]}")
ctx.close()
