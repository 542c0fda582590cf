#!/usr/bin/env python3
"""Checked timing of standalone window-table MSMs (uniform scalars), best of 5 after warm-up -- knob A/B helper.
usage: msm_time.py GROUP LOG_N [LOG_N ...]   (HK_FOLD_S / HK_REDUCE_G / HK_ACCUM_S from the environment)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from hekaton_system_b200 import codec
from hekaton_system_b200.backend import Context
from sweep import best_ms, limb_dot_mod

group = int(sys.argv[1])
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = Context("bls12_381", 0, stream=stream.cuda_stream)
cv = ctx.curve
L = cv.fq_limbs
rng = np.random.default_rng(3)
gen = (codec.g1_affine_to_limbs([cv.g1], cv) if group == 1 else codec.g2_affine_to_limbs([cv.g2], cv))[0]
out = []
for lg in map(int, sys.argv[2:]):
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
    ms = best_ms(stream, lambda: ctx.msm(group, h, d_sc.data_ptr(), n=n, out=d_out.data_ptr()))
    jac = d_out.cpu().numpy().view(np.uint64)[: 3 * (L if group == 1 else 2 * L)].reshape(1, -1)
    aff, inf = ctx.into_affine(group, jac)
    ok = bool(np.array_equal(aff[0], exp) and not inf[0])
    out.append(f"2^{lg}: {ms:.3f} ms {'exact' if ok else 'WRONG'}")
    ctx.bases_free(h)
    ctx.bases_free(he)
print(f"G{group} FOLD_S={os.environ.get('HK_FOLD_S', '-')} REDUCE_G={os.environ.get('HK_REDUCE_G', '-')}: " + "  ".join(out))
