"""Error behaviour at the C ABI: malformed inputs are rejected on the host (negative code + message) or made
harmless on the device -- never an out-of-bounds access.  (compute-sanitizer is not available on the GPU pool, so these
argument checks are the memory-safety evidence)"""
import ctypes as C

import numpy as np
import pytest

from hekaton_system_b200 import backend as B
from hekaton_system_b200 import codec
from hekaton_system_b200 import cp_groth16 as host
from oracle import cpgroth16 as og
from oracle.ec import groups

pytestmark = pytest.mark.gpu


def test_r1cs_upload_rejects_malformed_csr(ctx):
    cv = ctx.curve
    one = codec.to_mont([1], cv.r, 4)
    good = (np.array([0, 1, 2], dtype=np.uint32), np.array([0, 1], dtype=np.uint32), np.concatenate([one, one]))
    h = ctx.r1cs_upload(2, 1, 2, [good, good, good])
    ctx.r1cs_free(h)
    bad_start = (np.array([1, 1, 2], dtype=np.uint32), good[1], good[2])            # row_ptr[0] != 0
    decreasing = (np.array([0, 2, 1], dtype=np.uint32), good[1], good[2])           # not monotone
    bad_col = (good[0], np.array([0, 7], dtype=np.uint32), good[2])                 # column out of range
    for bad in (bad_start, decreasing, bad_col):
        with pytest.raises(B.HkError) as e:
            ctx.r1cs_upload(2, 1, 2, [good, bad, good])
        assert e.value.code == -1


def test_async_calls_need_flag_buffers(ctx):
    """HK_HOST_PINNED_ASYNC writes the infinity flags after the call returns: a NULL flag pointer is an error
    (a stack scratch would dangle), for hk_commit and hk_prove alike."""
    cv = ctx.curve
    circ = og.synthetic_circuit(cv.name, 30, stage0_witnesses=3, seed=5)
    td = og.sample_trapdoor(cv.name, 2, 6)
    opk = og.generate_parameters(cv.name, circ, td)
    pk = host.ProvingKey.from_key(ctx, opk, circ)
    z = codec.to_mont(circ.full_assignment(), cv.r, 4)
    one = codec.to_mont([1], cv.r, 4)
    out = np.zeros(4 * ctx.L, dtype=np.uint64)
    lib = ctx.lib
    rc = lib.hk_commit(ctx.h, C.c_uint64(pk.handle), C.c_uint32(0), C.c_void_p(z.ctypes.data), C.c_void_p(one.ctypes.data),
                       C.c_uint32(B.HK_HOST_PINNED_ASYNC), C.c_void_p(out.ctypes.data), None)
    assert rc == -1
    rc = lib.hk_prove(ctx.h, C.c_uint64(pk.handle), C.c_void_p(z.ctypes.data), C.c_void_p(one.ctypes.data),
                      C.c_void_p(one.ctypes.data), C.c_void_p(one.ctypes.data), C.c_uint32(B.HK_HOST_PINNED_ASYNC),
                      C.c_void_p(out.ctypes.data), C.c_void_p(out.ctypes.data), C.c_void_p(out.ctypes.data), None)
    assert rc == -1
    # the blocking memspaces accept NULL flags (the scratch lives until the call has synchronised)
    rc = lib.hk_commit(ctx.h, C.c_uint64(pk.handle), C.c_uint32(0), C.c_void_p(z.ctypes.data), C.c_void_p(one.ctypes.data),
                       C.c_uint32(B.HK_HOST), C.c_void_p(out.ctypes.data), None)
    assert rc == 0


def test_non_canonical_scalars_cannot_corrupt_memory(ctx):
    """HK_SCALAR_BIGINT scalars are canonical by contract; values with bits above the field size used to index past
    the bucket array.  Now the top digit is clamped: the call completes (with a meaningless sum) and the context keeps
    producing exact results."""
    cv = ctx.curve
    G = groups(cv.name)[0]
    n = 300
    pts = [G.mul(G.gen, 3 + i) for i in range(n)]
    limbs, inf = codec.g1_affine_to_limbs(pts, cv)
    h = ctx.upload_g1(limbs, inf)
    for table in (False, True):
        if table:
            ctx.bases_precompute(h)
        junk = np.full((n, 4), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        ctx.msm(1, h, junk)                                                  # must not fault
        sc = [(7 * i + 1) % cv.r for i in range(n)]
        aff, oinf = ctx.into_affine(1, ctx.msm(1, h, codec.ints_to_limbs(sc, 4)))
        assert codec.g1_affine_from_limbs(aff, cv, oinf)[0] == G.msm_naive(pts, sc)
    ctx.bases_free(h)


def test_unknown_option_and_bad_values(ctx):
    with pytest.raises(B.HkError):
        ctx.set_option(99, 1)
    with pytest.raises(B.HkError):
        ctx.set_option(B.OPT_FOLD_SEG, 5)
    with pytest.raises(B.HkError):
        ctx.set_option(B.OPT_REDUCE_CHUNK, 3)
    ctx.set_option(B.OPT_FOLD_SEG, 0)
