"""GPU parity: hk_msm_g1 / hk_msm_g2 vs the oracle (bit-exact after affine normalisation), with
the edge cases the reference's call sites exercise (zip-to-shorter, empty, 0/1-heavy scalars)."""
import random

import numpy as np
import pytest

from hekaton_system_b200 import codec
from hekaton_system_b200 import backend as B
from oracle.ec import groups

pytestmark = pytest.mark.gpu


def _bases(G, n, rng, r):
    return [G.mul(G.gen, rng.randrange(1, r)) for _ in range(n)]


def _upload(ctx, group, pts):
    cv = ctx.curve
    if group == 1:
        return ctx.upload_g1(*codec.g1_affine_to_limbs(pts, cv))
    return ctx.upload_g2(*codec.g2_affine_to_limbs(pts, cv))


def _msm_affine(ctx, group, handle, scalars, **kw):
    cv = ctx.curve
    sc = codec.ints_to_limbs(scalars, 4) if scalars else np.zeros((0, 4), dtype=np.uint64)
    jac = ctx.msm(group, handle, sc, n=len(scalars), **kw)
    aff, inf = ctx.into_affine(group, jac)
    dec = codec.g1_affine_from_limbs if group == 1 else codec.g2_affine_from_limbs
    return dec(aff, cv, inf)[0]


@pytest.mark.parametrize("table", [False, True], ids=["plain", "window_table"])
@pytest.mark.parametrize("group", [1, 2])
def test_msm_small_and_edges(ctx, group, table):
    cv = ctx.curve
    G = groups(cv.name)[group - 1]
    rng = random.Random(group)
    n = 70 if group == 1 else 24
    pts = _bases(G, n, rng, cv.r)
    pts[4] = None                       # base at infinity
    pts[9] = pts[8]                     # duplicate base  (P + P inside a bucket)
    pts[11] = G.neg(pts[10])            # P and -P        (cancellation inside a bucket)
    h = _upload(ctx, group, pts)
    if table:
        ctx.bases_precompute(h)         # window-shifted table: one bucket set, no Horner tail
    uni = [rng.randrange(cv.r) for _ in range(n)]
    same = list(uni)
    same[9] = same[8]
    same[11] = same[10]
    cases = {
        "uniform": uni,
        "dup_and_cancel": same,
        "zeros": [0] * n,
        "ones": [1] * n,
        "r_minus_1": [cv.r - 1] * n,
        "bits": [rng.randrange(2) for _ in range(n)],
        "small": [rng.randrange(256) for _ in range(n)],
        "single_nonzero": [0] * (n - 1) + [12345],
        "top_bit": [(1 << (cv.fr_bits - 1)) + 1] * n,
    }
    for name, sc in cases.items():
        assert _msm_affine(ctx, group, h, sc) == G.msm_naive(pts, sc), name
    # sizes around the window-size switch and n = 0, 1, 2
    for m in (0, 1, 2, 31, 32, 33):
        if m <= n:
            assert _msm_affine(ctx, group, h, uni[:m]) == G.msm_naive(pts[:m], uni[:m]), m
    # msm_bigint zips to the shorter (committer.rs:110-113): more scalars than bases, and an offset
    assert _msm_affine(ctx, group, h, uni + [5, 6, 7]) == G.msm_naive(pts, uni)
    assert _msm_affine(ctx, group, h, uni[:10], offset=3) == G.msm_naive(pts[3:13], uni[:10])
    assert _msm_affine(ctx, group, h, uni[:10], offset=n - 2) == G.msm_naive(pts[n - 2:], uni[:2])
    # Montgomery-form scalars (E::G1::msm with Fr inputs; prover.rs:117,129)
    sm = codec.to_mont(uni, cv.r, 4)
    jac = ctx.msm(group, h, sm, repr_=B.SCALAR_MONT)
    aff, inf = ctx.into_affine(group, jac)
    dec = codec.g1_affine_from_limbs if group == 1 else codec.g2_affine_from_limbs
    assert dec(aff, cv, inf)[0] == G.msm_naive(pts, uni)
    ctx.bases_free(h)


def test_msm_checked_length_error(ctx):
    """VariableBaseMSM::msm returns Err when lengths differ (committer.rs:89 unwraps it)."""
    cv = ctx.curve
    G = groups(cv.name)[0]
    pts = _bases(G, 4, random.Random(1), cv.r)
    h = _upload(ctx, 1, pts)
    sc = codec.ints_to_limbs([1, 2, 3], 4)
    with pytest.raises(B.HkError) as e:
        ctx.msm_checked(h, 4, sc)
    assert e.value.code == -4
    out = ctx.msm_checked(h, 3, sc)
    aff, inf = ctx.into_affine(1, out)
    assert codec.g1_affine_from_limbs(aff, cv, inf)[0] == G.msm_naive(pts[:3], [1, 2, 3])
    with pytest.raises(B.HkError):
        ctx.msm(2, h, sc)            # wrong group
    with pytest.raises(B.HkError):
        ctx.msm(1, 987654, sc)       # unknown handle


def test_fixed_base_generation_matches_oracle(ctx):
    cv = ctx.curve
    G1, G2 = groups(cv.name)
    rng = random.Random(5)
    ks = [0, 1, 2, cv.r - 1, 15, 16, 1 << 200] + [rng.randrange(cv.r) for _ in range(20)]
    for group, G, enc, dec in ((1, G1, codec.g1_affine_to_limbs, codec.g1_affine_from_limbs),
                               (2, G2, codec.g2_affine_to_limbs, codec.g2_affine_from_limbs)):
        gen = enc([G.gen], cv)[0]
        h = ctx.bases_from_scalars(group, gen, codec.ints_to_limbs(ks, 4))
        got = dec(ctx.bases_download(h, group), cv)
        assert got == [G.mul(G.gen, k) for k in ks]
        ctx.bases_free(h)


@pytest.mark.parametrize("table", [False, True], ids=["plain", "window_table"])
@pytest.mark.parametrize("group,n", [(1, 3000), (1, 40000), (2, 3000)])
def test_msm_medium_vs_exponent_oracle(ctx, group, n, table):
    """Bases are k_i * G generated on the device (checked above), so the expected result is the
    single scalar multiple (sum k_i s_i) * G -- exact, independent of any MSM algorithm."""
    cv = ctx.curve
    G = groups(cv.name)[group - 1]
    rng = random.Random(n + group)
    ks = [rng.randrange(1, cv.r) for _ in range(n)]
    enc = codec.g1_affine_to_limbs if group == 1 else codec.g2_affine_to_limbs
    h = ctx.bases_from_scalars(group, enc([G.gen], cv)[0], codec.ints_to_limbs(ks, 4))
    if table:
        ctx.bases_precompute(h)
    for kind in ("uniform", "witness_like", "all_same"):
        if kind == "uniform":
            sc = [rng.randrange(cv.r) for _ in range(n)]
        elif kind == "witness_like":      # 90 % in {0,1}: SHA-256-style witnesses
            sc = [rng.randrange(2) if rng.random() < 0.9 else rng.randrange(cv.r) for _ in range(n)]
        else:
            sc = [0xdeadbeefcafe] * n
        e = sum(k * s for k, s in zip(ks, sc)) % cv.r
        assert _msm_affine(ctx, group, h, sc) == G.mul(G.gen, e), kind
    ctx.bases_free(h)


def test_msm_batch(ctx):
    cv = ctx.curve
    G = groups(cv.name)[0]
    rng = random.Random(77)
    n, K = 500, 3
    ks = [rng.randrange(1, cv.r) for _ in range(n)]
    h = ctx.bases_from_scalars(1, codec.g1_affine_to_limbs([G.gen], cv)[0], codec.ints_to_limbs(ks, 4))
    scs = [[rng.randrange(cv.r) for _ in range(n)] for _ in range(K)]
    jac = ctx.msm(1, h, np.concatenate([codec.ints_to_limbs(s, 4) for s in scs]), n=n, K=K)
    aff, inf = ctx.into_affine(1, jac)
    got = codec.g1_affine_from_limbs(aff, cv, inf)
    for k in range(K):
        e = sum(a * b for a, b in zip(ks, scs[k])) % cv.r
        assert got[k] == G.mul(G.gen, e)


def test_forked_contexts_share_resources(ctx):
    """hk_ctx_fork: same handles, own stream + scratch; results identical from either context,
    also when driven from two host threads at once (node.rs:745-795 threading model)."""
    import threading
    cv = ctx.curve
    G = groups(cv.name)[0]
    rng = random.Random(9)
    n = 2000
    ks = [rng.randrange(1, cv.r) for _ in range(n)]
    h = ctx.bases_from_scalars(1, codec.g1_affine_to_limbs([G.gen], cv)[0], codec.ints_to_limbs(ks, 4))
    ctx.bases_precompute(h)
    forks = [ctx.fork() for _ in range(3)]
    scs = [[rng.randrange(cv.r) for _ in range(n)] for _ in forks]
    exp = [G.mul(G.gen, sum(a * b for a, b in zip(ks, sc)) % cv.r) for sc in scs]
    got = [None] * len(forks)

    def work(i):
        for _ in range(3):
            got[i] = _msm_affine(forks[i], 1, h, scs[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(forks))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert got == exp
    assert _msm_affine(ctx, 1, h, scs[0]) == exp[0]
    for f in forks:
        f.close()
    ctx.bases_free(h)


@pytest.mark.parametrize("group,log_n", [(1, 20), (2, 18)])
def test_msm_full_size_exact(ctx, group, log_n):
    """BASELINE.json's headline size (G1 MSM at 2^20; G2 at 2^18), window-table mode, checked EXACTLY:
    bases are k_i * G generated on the device, so the expected result is (sum k_i s_i mod r) * G."""
    cv = ctx.curve
    if cv.name != "bls12_381":
        pytest.skip("full-size run on the primary curve")
    G = groups(cv.name)[group - 1]
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    ks[:, 3] >>= 2
    enc = codec.g1_affine_to_limbs if group == 1 else codec.g2_affine_to_limbs
    h = ctx.bases_from_scalars(group, enc([G.gen], cv)[0], ks)
    ctx.bases_precompute(h, 1.0)
    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    sc[:, 3] >>= 2
    sc[::7] = 0
    sc[1::7] = 0
    sc[1::7, 0] = 1                      # some zeros and ones, like witness vectors
    jac = ctx.msm(group, h, sc)
    aff, inf = ctx.into_affine(group, jac)
    dec = codec.g1_affine_from_limbs if group == 1 else codec.g2_affine_from_limbs
    ki = codec.limbs_to_ints(ks)
    si = codec.limbs_to_ints(sc)
    e = sum(a * b for a, b in zip(ki, si)) % cv.r
    assert dec(aff, cv, inf)[0] == G.mul(G.gen, e)
    ctx.bases_free(h)


@pytest.mark.parametrize("group", [1, 2])
def test_scalar_mul_matches_oracle(ctx, group):
    """hk_scalar_mul_* = the reference's `scalar_pairing` (pairing_ops.rs:32-39): out[i] = s_i * P_i, normalised.
    Zero / one / r - 1 scalars, the point at infinity, canonical and Montgomery scalars."""
    cv = ctx.curve
    G = groups(cv.name)[group - 1]
    rng = random.Random(60 + group)
    enc, dec = ((codec.g1_affine_to_limbs, codec.g1_affine_from_limbs) if group == 1 else
                (codec.g2_affine_to_limbs, codec.g2_affine_from_limbs))
    pts = [G.mul(G.gen, rng.randrange(1, cv.r)) for _ in range(9)] + [None, G.gen]
    sc = [0, 1, 2, cv.r - 1, 1 << 254 if cv.r > (1 << 254) else 1 << 253, 0xffffffff] + \
         [rng.randrange(cv.r) for _ in range(3)] + [rng.randrange(cv.r), cv.r - 2]
    limbs, inf = enc(pts, cv)
    exp = [None if p is None else G.mul(p, s) for p, s in zip(pts, sc)]
    out, oinf = ctx.scalar_mul(group, limbs, codec.ints_to_limbs(sc, 4), inf=inf)
    assert dec(out, cv, oinf) == exp
    out, oinf = ctx.scalar_mul(group, limbs, codec.to_mont(sc, cv.r, 4), inf=inf, repr_=B.SCALAR_MONT)
    assert dec(out, cv, oinf) == exp
    # twist-power pattern of the aggregator: P_i * r^i (aggregation.rs:225-242)
    tw = rng.randrange(cv.r)
    pw = [pow(tw, i, cv.r) for i in range(len(pts))]
    out, oinf = ctx.scalar_mul(group, limbs, codec.ints_to_limbs(pw, 4), inf=inf)
    assert dec(out, cv, oinf) == [None if p is None else G.mul(p, s) for p, s in zip(pts, pw)]


@pytest.mark.parametrize("opts", [
    {B.OPT_FOLD_SEG: 4, B.OPT_ACCUM_SEG_MIN: 1},                      # one item per thread: every run goes through the folds
    {B.OPT_ACCUM_WAVES: 1, B.OPT_ACCUM_SEG_MIN: 2, B.OPT_FOLD_SEG: 32},   # grid capped at one wave: device-side segment length
    {B.OPT_ACCUM_SEG_MIN: 512, B.OPT_REDUCE_CHUNK: 16},               # a handful of threads own the whole list
    {B.OPT_ACCUM_WAVES: 64, B.OPT_ACCUM_SEG_MIN: 3, B.OPT_FOLD_SEG: 8, B.OPT_REDUCE_CHUNK: 1},
    {B.OPT_SORT_GLOBAL_ATOMICS: 1},                                   # table sets through the global-atomic sort
    {B.OPT_G2_LANE_PAIR: 1, B.OPT_ACCUM_SEG_MIN: 5},                  # Fq2 values spread over lane pairs (BLS12-381 G2)
    {B.OPT_G2_LANE_PAIR: 2, B.OPT_ACCUM_SEG_MIN: 7},                  # ... and never: one thread per G2 addition
    {B.OPT_QUAD_MAX_THREADS: 1, B.OPT_ACCUM_SEG_MIN: 2},              # no quad-cooperative tail kernels at all
    {B.OPT_QUAD_MAX_THREADS: 1 << 24, B.OPT_ACCUM_SEG_MIN: 2, B.OPT_FOLD_SEG: 8},   # every tail kernel quad-cooperative
    {B.OPT_QUAD_MAX_THREADS: 1 << 24, B.OPT_NO_FOLD_RESOLVE: 1, B.OPT_ACCUM_SEG_MIN: 1},   # one launch per fold level to the end
    {B.OPT_QUAD_MAX_THREADS: 1 << 24, B.OPT_ACCUM_SEG_MIN: 64, B.OPT_FOLD_SEG: 4},   # borders closed at once from the first level on
])
def test_tuning_options_never_change_results(ctx, opts):
    """hk_ctx_set_option: the accumulation geometry (waves, minimum segment) and the fold / reduction chunking only
    move work between kernels.  Forced to extreme values on a forked context and run over
    the cases that stress the segment borders: runs of every length 0..40 in one bucket, P + P, P - P, bases at
    infinity, uniform and witness-like scalars, G1 / G2, table and plain mode."""
    cv = ctx.curve
    f = ctx.fork()
    for k, v in opts.items():
        f.set_option(k, v)
    rng = random.Random(900 + sum(opts.values()))
    try:
        for group in (1, 2):
            G = groups(cv.name)[group - 1]
            n = 700 if group == 1 else 200
            pts = [G.mul(G.gen, rng.randrange(1, cv.r)) for _ in range(n // 4)]
            pts = pts + pts[: n // 8] + [G.neg(p) for p in pts[: n // 8]] + [None] * 5      # duplicates, negations, infinity
            while len(pts) < n:
                pts.append(G.mul(G.gen, rng.randrange(1, cv.r)))
            rng.shuffle(pts)
            h = _upload(f, group, pts)
            for table in (False, True):
                if table:
                    f.bases_precompute(h)
                for kind in range(4):
                    if kind == 0:
                        sc = [rng.randrange(cv.r) for _ in range(n)]
                    elif kind == 1:          # every scalar small: a few buckets with long runs of every length
                        sc = [rng.randrange(0, 6) for _ in range(n)]
                    elif kind == 2:          # one bucket only
                        sc = [1] * n
                    else:
                        sc = [rng.randrange(2) if rng.random() < 0.9 else rng.randrange(cv.r) for _ in range(n)]
                    assert _msm_affine(f, group, h, sc) == G.msm_naive(pts, sc), (group, table, kind)
            # runs of length 0..40 of the same point in bucket 1: sum = L * P
            for L_ in (0, 1, 2, 3, 7, 8, 9, 31, 32, 33, 40):
                sc = [1] * L_ + [0] * (n - L_)
                acc = None
                for p_ in pts[:L_]:
                    if p_ is not None:
                        acc = G.add(acc, p_)
                assert _msm_affine(f, group, h, sc) == acc, (group, L_)
            f.bases_free(h)
    finally:
        f.close()
