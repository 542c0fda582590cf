"""oracle/pairing.py: the pairing is pinned by the properties that define it (bilinear in both arguments,
non-degenerate, values of order r) on the public standard generators, and verify_proof (verifier.rs:23-71) accepts
the oracle's own setup -> commit -> prove output and rejects tampered proofs -- the reference's acceptance test
(cp-groth16/src/lib.rs:139-180, 273-313) restated."""
import random

import pytest

from oracle import cpgroth16 as og
from oracle import pairing as pr
from oracle.ec import groups
from oracle.params import get_curve


@pytest.fixture(scope="module", params=["bls12_381", "bn254"])
def name(request):
    return request.param


def test_pairing_properties(name):
    c = get_curve(name)
    G1, G2 = groups(name)
    one = pr.gt_one(name)
    e = pr.pairing(name, G1.gen, G2.gen)
    assert e != one                                             # non-degenerate
    assert pr.gt_pow(name, e, c.r) == one                       # GT has order r
    rng = random.Random(5)
    a, b = rng.randrange(1, c.r), rng.randrange(1, c.r)
    P, Q = G1.mul(G1.gen, a), G2.mul(G2.gen, b)
    assert pr.pairing(name, P, Q) == pr.gt_pow(name, e, a * b % c.r)          # e(aP, bQ) = e(P, Q)^(ab)
    assert pr.pairing(name, P, G2.gen) == pr.pairing(name, G1.gen, G2.mul(G2.gen, a))
    # additivity in each argument through the multi Miller loop
    P2 = G1.mul(G1.gen, b)
    lhs = pr.final_exponentiation(name, pr.multi_miller_loop(name, [P, P2], [Q, Q]))
    assert lhs == pr.pairing(name, G1.add(P, P2), Q)
    assert pr.final_exponentiation(name, pr.multi_miller_loop(name, [P, G1.neg(P)], [Q, Q])) == one
    assert pr.pairing(name, None, Q) == one and pr.pairing(name, P, None) == one


@pytest.mark.parametrize("stage0", [0, 5])
def test_verify_proof_accepts_oracle_proofs_and_rejects_tampering(name, stage0):
    c = get_curve(name)
    G1, G2 = groups(name)
    circ = og.synthetic_circuit(name, 40, stage0_witnesses=stage0, seed=3)
    if stage0 == 0:
        circ.stage_ranges = [(0, len(circ.witness))]
    ns = len(circ.stage_ranges)
    td = og.sample_trapdoor(name, ns, 4)
    pk = og.generate_parameters(name, circ, td)
    kappas = [0x1234567 + i for i in range(ns - 1)]
    ds = [og.commit(pk, circ, i, k) for i, k in enumerate(kappas)]
    A, B, C = og.prove(pk, circ, 0xabcdef, 0x13579b, kappas)
    pvk = pr.prepare_verifying_key(pk)
    pub = circ.instance[1:]
    assert pr.verify_proof(pvk, A, B, C, ds, pub)
    assert not pr.verify_proof(pvk, G1.add(A, G1.gen), B, C, ds, pub)
    assert not pr.verify_proof(pvk, A, B, G1.add(C, G1.gen), ds, pub)
    assert not pr.verify_proof(pvk, A, G2.add(B, G2.gen), C, ds, pub)
    bad_pub = list(pub)
    bad_pub[0] = (bad_pub[0] + 1) % c.r
    assert not pr.verify_proof(pvk, A, B, C, ds, bad_pub)
    if ds:
        assert not pr.verify_proof(pvk, A, B, C, [G1.add(ds[0], G1.gen)], pub)
    with pytest.raises(ValueError):
        pr.verify_proof(pvk, A, B, C, ds, pub[:-1])              # MalformedVerifyingKey (verifier.rs:51-53)
    # the trapdoor known answer passes the independent pairing check as well
    exp = og.trapdoor_expected(pk, circ, td, 0xabcdef, 0x13579b, kappas)
    assert pr.verify_proof(pvk, exp["A"], exp["B"], exp["C"], exp["D"], pub)
