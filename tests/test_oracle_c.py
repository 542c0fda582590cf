"""Pins the C restatement (oracle/c, the CPU baseline) against the Python big-integer oracle."""
import random

import numpy as np
import pytest

from hekaton_system_b200 import codec
from hekaton_system_b200.cp_groth16 import StagedAssignment
from hekaton_system_b200.curves import get_curve
from oracle import cref, poly, cpgroth16 as og
from oracle.ec import groups
from oracle.params import BLS12_381, BN254


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_c_field_mul(c):
    rng = random.Random(1)
    for field, mod, L in ((0, c.r, 4), (1, c.p, c.fq_limbs64)):
        R = 1 << (64 * L)
        a = [0, 1, mod - 1, R % mod] + [rng.randrange(mod) for _ in range(200)]
        b = [mod - 1, mod - 1, mod - 1, R % mod] + [rng.randrange(mod) for _ in range(200)]
        got = codec.limbs_to_ints(cref.field_mul(c.cid, field, codec.ints_to_limbs(a, L), codec.ints_to_limbs(b, L)))
        assert got == [x * y * pow(R, -1, mod) % mod for x, y in zip(a, b)]


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
@pytest.mark.parametrize("threads", [1, 4])
def test_c_ntt(c, threads):
    rng = random.Random(2)
    for log_n in (0, 1, 4, 9):
        n = 1 << log_n
        x = [rng.randrange(c.r) for _ in range(n)]
        dom = poly.Domain(c, n)
        co = dom.get_coset(c.fr_generator)
        for inverse, coset, ref in ((False, False, dom.fft), (True, False, dom.ifft), (False, True, co.fft),
                                    (True, True, co.ifft)):
            buf = codec.to_mont(x, c.r, 4)
            cref.ntt(c.cid, buf, log_n, inverse, coset, threads)
            assert codec.from_mont(buf, c.r) == ref(x)


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_c_msm(c):
    cv = get_curve(c.name)
    rng = random.Random(3)
    G1, G2 = groups(c)
    for group, G, n, enc, dec in ((1, G1, 45, codec.g1_affine_to_limbs, codec.g1_affine_from_limbs),
                                  (2, G2, 34, codec.g2_affine_to_limbs, codec.g2_affine_from_limbs)):
        pts = [G.mul(G.gen, rng.randrange(1, c.r)) for _ in range(n)]
        pts[2] = None
        pts[5] = pts[4]
        sc = [rng.randrange(c.r) for _ in range(n)]
        sc[0], sc[1], sc[5] = 0, 1, sc[4]
        arr, inf = enc(pts, cv)
        for m in (0, 1, 31, n):
            out, oinf = cref.msm(c.cid, group, arr[:m], inf[:m], codec.ints_to_limbs(sc[:m], 4) if m else np.zeros((0, 4), dtype=np.uint64), threads=3)
            got = None if oinf else dec(out, cv)[0]
            assert got == G.msm_naive(pts[:m], sc[:m]), (group, m)
        out, oinf = cref.msm(c.cid, group, arr, inf, codec.to_mont(sc, c.r, 4), repr_=1, threads=2)
        assert dec(out, cv)[0] == G.msm_naive(pts, sc)


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_c_witness_map_and_prove(c):
    cv = get_curve(c.name)
    circ = og.synthetic_circuit(c, 30, stage0_witnesses=5, seed=12)
    staged = StagedAssignment.from_circuit(circ, cv)
    z = circ.full_assignment()
    zm = codec.to_mont(z, c.r, 4)
    h = cref.witness_map(c.cid, 30, staged.l, staged.csr(), zm, threads=2)
    assert codec.from_mont(h, c.r) == poly.witness_map(c, circ.r1cs, z)
    td = og.sample_trapdoor(c, 2, 13)
    opk = og.generate_parameters(c, circ, td)
    e1 = lambda pts: codec.g1_affine_to_limbs(pts, cv)[0]
    e2 = lambda pts: codec.g2_affine_to_limbs(pts, cv)[0]
    starts = [s for s, _ in circ.stage_ranges] + [circ.stage_ranges[-1][1]]
    cp = cref.CpuProver(c.cid, 30, staged.l, len(circ.witness), staged.csr(), starts, e1(opk.a_g), e1(opk.b_g),
                        e2(opk.b_h), e1(opk.h_g), [e1(v) for v in opk.deltas_abc_g], e1(opk.deltas_g),
                        e1([opk.alpha_g]), e1([opk.beta_g]), e2([opk.beta_h]), e2([opk.last_delta_h]))
    kappa, r_, s_ = 0x77777, 0x123456789abcdef, 0xfedcba
    fr = lambda v: codec.to_mont([v], c.r, 4)[0]
    a, b, cc = cp.prove(zm, fr(r_), fr(s_), codec.to_mont([kappa], c.r, 4), threads=2)
    exp = og.trapdoor_expected(opk, circ, td, r_, s_, [kappa])
    assert codec.g1_affine_from_limbs(a, cv)[0] == exp["A"]
    assert codec.g2_affine_from_limbs(b, cv)[0] == exp["B"]
    assert codec.g1_affine_from_limbs(cc, cv)[0] == exp["C"]
    # r == 0 branch (prover.rs:92)
    a, b, cc = cp.prove(zm, fr(0), fr(s_), codec.to_mont([kappa], c.r, 4))
    exp = og.trapdoor_expected(opk, circ, td, 0, s_, [kappa])
    assert (codec.g1_affine_from_limbs(a, cv)[0], codec.g1_affine_from_limbs(cc, cv)[0]) == (exp["A"], exp["C"])
