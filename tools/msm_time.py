#!/usr/bin/env python3
"""Checked timing of standalone window-table MSMs (uniform scalars), best of 5 after warm-up -- knob A/B helper.
usage: msm_time.py GROUP LOG_N[,LOG_N...] [OPTSET ...]
  OPTSET = comma-separated hk_ctx_set_option settings, e.g.  waves=2,smin=64  fold=16  (names: waves smin fold chunk); every OPTSET is timed on a fork of the same context (same bases, same scalars); "-" = defaults."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from hekaton_system_b200 import backend as B
from hekaton_system_b200 import codec
from hekaton_system_b200.backend import Context
from sweep import best_ms, limb_dot_mod

from msm_time_names import NAMES

group = int(sys.argv[1])
sizes = [int(x) for x in sys.argv[2].split(",")]
optsets = sys.argv[3:] or ["-"]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = Context("bls12_381", 0, stream=stream.cuda_stream)
cv = ctx.curve
L = cv.fq_limbs
rng = np.random.default_rng(3)
gen = (codec.g1_affine_to_limbs([cv.g1], cv) if group == 1 else codec.g2_affine_to_limbs([cv.g2], cv))[0]
rows = {o: [] for o in optsets}
for lg in sizes:
    n = 1 << lg
    ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    ks[:, 3] >>= 2
    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    sc[:, 3] >>= 2
    h = ctx.bases_from_scalars(group, gen, ks)
    ctx.bases_precompute(h, 1.0)
    he = ctx.bases_from_scalars(group, gen, codec.ints_to_limbs([limb_dot_mod(ks, sc, cv.r)], 4))
    exp = ctx.bases_download(he, group)[0]
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(128, dtype=torch.int64, device="cuda")
    for o in optsets:
        f = ctx.fork(stream=stream.cuda_stream)
        if o != "-":
            for kv in o.split(","):
                k, v = kv.split("=")
                f.set_option(NAMES[k], int(v))
        ms = best_ms(stream, lambda: f.msm(group, h, d_sc.data_ptr(), n=n, out=d_out.data_ptr()), reps=5)
        jac = d_out.cpu().numpy().view(np.uint64)[: 3 * (L if group == 1 else 2 * L)].reshape(1, -1)
        aff, inf = ctx.into_affine(group, jac)
        ok = bool(np.array_equal(aff[0], exp) and not inf[0])
        rows[o].append(f"2^{lg}: {ms:.3f} ms {'exact' if ok else 'WRONG'}")
        f.close()
    ctx.bases_free(h)
    ctx.bases_free(he)
for o in optsets:
    print(f"G{group} [{o}]: " + "  ".join(rows[o]))
