"""GPU parity: hk_ntt / hk_qap_h / hk_qap_witness_map vs the oracle's Radix2EvaluationDomain and
LibsnarkReduction restatement (bit-exact), plus size-independent properties at large sizes."""
import random

import numpy as np
import pytest

from hekaton_system_b200 import codec
from oracle import poly
from oracle import cpgroth16 as og

pytestmark = pytest.mark.gpu


def _rand(rng, n, r):
    return [rng.randrange(r) for _ in range(n)]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10, 11, 12, 13])
def test_ntt_matches_oracle(ctx, log_n):
    cv = ctx.curve
    rng = random.Random(log_n)
    n = 1 << log_n
    x = _rand(rng, n, cv.r)
    if log_n >= 3:
        x[0], x[1], x[2] = 0, 1, cv.r - 1
    dom = poly.Domain(cv.name, n)
    co = dom.get_coset(cv.fr_generator)
    for inverse, coset, ref in ((False, False, dom.fft), (True, False, dom.ifft), (False, True, co.fft),
                                (True, True, co.ifft)):
        buf = codec.to_mont(x, cv.r, 4)
        ctx.ntt(buf, log_n, inverse=inverse, coset=coset)
        assert codec.from_mont(buf, cv.r) == ref(x), (log_n, inverse, coset)


def test_ntt_batch_and_device_pointer(ctx):
    import torch
    cv = ctx.curve
    rng = random.Random(3)
    log_n, K = 9, 3
    xs = [_rand(rng, 1 << log_n, cv.r) for _ in range(K)]
    buf = np.concatenate([codec.to_mont(x, cv.r, 4) for x in xs])
    ctx.ntt(buf, log_n, K=K)
    dom = poly.Domain(cv.name, 1 << log_n)
    for k in range(K):
        assert codec.from_mont(buf[k << log_n:(k + 1) << log_n], cv.r) == dom.fft(xs[k])
    # device-resident in/out
    t = torch.from_numpy(codec.to_mont(xs[0], cv.r, 4).view(np.int64)).cuda()
    ctx.ntt(t.data_ptr(), log_n)
    ctx.sync()
    assert codec.from_mont(t.cpu().numpy().view(np.uint64), cv.r) == dom.fft(xs[0])


@pytest.mark.parametrize("log_n", [16, 20, 21, 22, 23])
def test_ntt_large_properties(ctx, log_n):
    """Round trip, linearity and a single-coefficient known answer at sizes the O(N log N) oracle
    does not reach in seconds."""
    cv = ctx.curve
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    a = rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)   # < r (both curves) as raw Montgomery residues
    orig = a.copy()
    for coset in (False, True):
        ctx.ntt(a, log_n, coset=coset)
        assert not np.array_equal(a, orig)
        ctx.ntt(a, log_n, inverse=True, coset=coset)
        assert np.array_equal(a, orig)
    # delta at index 1 -> powers of the domain generator: out[k] = w^k
    d = np.zeros((n, 4), dtype=np.uint64)
    d[1] = codec.to_mont([1], cv.r, 4)[0]
    ctx.ntt(d, log_n)
    dom = poly.Domain(cv.name, n)
    ks = [0, 1, 2, 3, n // 2, n - 1, 12345 % n]
    got = codec.from_mont(d[ks], cv.r)
    assert got == [pow(dom.group_gen, k, cv.r) for k in ks]


@pytest.mark.parametrize("log_n", [1, 4, 9, 12])
def test_qap_h_matches_oracle(ctx, log_n):
    cv = ctx.curve
    rng = random.Random(40 + log_n)
    n = 1 << log_n
    a, b, c = (_rand(rng, n, cv.r) for _ in range(3))
    dom = poly.Domain(cv.name, n)
    ref = poly.h_from_evals(cv.name, dom, a, b, c)
    h = ctx.qap_h(codec.to_mont(a, cv.r, 4), codec.to_mont(b, cv.r, 4), codec.to_mont(c, cv.r, 4), log_n)
    assert codec.from_mont(h, cv.r) == ref


@pytest.mark.parametrize("log_n", [13, 21])
def test_qap_h_equals_composition_of_transforms(ctx, log_n):
    """hk_qap_h (DIF -> DIT -> pointwise -> DIF pipeline that never materialises a bit reversal; at
    log_n >= 21 its last pass scatters through top-passenger tiles) against the same h assembled from
    standalone hk_ntt calls (DIT with bit-reversed loads) and element-wise field ops: two disjoint kernel
    paths.  log_n = 21 is the VM shape's domain (BASELINE config 5)."""
    from hekaton_system_b200 import backend as B
    cv = ctx.curve
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    a, b, c = (rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64) for _ in range(3))   # Montgomery residues < r
    h = ctx.qap_h(a.copy(), b.copy(), c.copy(), log_n)
    ev = []
    for x in (a, b, c):
        y = x.copy()
        ctx.ntt(y, log_n, inverse=True)
        ctx.ntt(y, log_n, coset=True)
        ev.append(y)
    ab = ctx.field_op(B.FIELD_FR, B.FOP_MUL, ev[0], ev[1])
    num = ctx.field_op(B.FIELD_FR, B.FOP_SUB, ab, ev[2])
    g = cv.fr_generator
    zinv = pow(pow(g, n, cv.r) - 1, -1, cv.r)
    zm = np.repeat(codec.to_mont([zinv], cv.r, 4), n, axis=0)
    q = ctx.field_op(B.FIELD_FR, B.FOP_MUL, num, zm)
    ctx.ntt(q, log_n, inverse=True, coset=True)
    assert np.array_equal(q, h)


@pytest.mark.parametrize("ncons", [3, 50, 700])
def test_witness_map_matches_oracle(ctx, ncons):
    from hekaton_system_b200.cp_groth16 import StagedAssignment
    cv = ctx.curve
    circ = og.synthetic_circuit(cv.name, ncons, stage0_witnesses=4, seed=ncons)
    staged = StagedAssignment.from_circuit(circ, cv)
    r1 = ctx.r1cs_upload(ncons, staged.l, len(staged.witness), staged.csr())
    z = circ.full_assignment()
    ref = poly.witness_map(cv.name, circ.r1cs, z)
    h = ctx.witness_map(r1, codec.to_mont(z, cv.r, 4), n_out=len(ref))
    got = codec.from_mont(h, cv.r)
    assert got == ref
    assert got[-1] == 0          # prover.rs:128-129 relies on deg h <= N-2
    ctx.r1cs_free(r1)
