"""Oracle prover == trapdoor known answer; single- and two-stage circuits,
mirroring the reference's poly_commit_test pair (cp-groth16/src/lib.rs:139-180, 273-313)."""
import pytest

from oracle.params import BLS12_381, BN254
from oracle import cpgroth16 as cg


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_two_stage_trapdoor_kat(c):
    circ = cg.synthetic_circuit(c, 24, stage0_witnesses=5, seed=3)
    td = cg.sample_trapdoor(c, 2, 7)
    pk = cg.generate_parameters(c, circ, td)
    kappa, r_, s_ = 0x1234567, 0x987654321, 0x555
    D0 = cg.commit(pk, circ, 0, kappa)
    A, B, C = cg.prove(pk, circ, r_, s_, [kappa])
    exp = cg.trapdoor_expected(pk, circ, td, r_, s_, [kappa])
    assert (A, B, C, D0) == (exp["A"], exp["B"], exp["C"], exp["D"][0])


def test_single_stage_and_no_zk():
    c = BN254
    circ = cg.synthetic_circuit(c, 12, stage0_witnesses=0, seed=5)
    circ.stage_ranges = [(0, len(circ.witness))]
    td = cg.sample_trapdoor(c, 1, 8)
    pk = cg.generate_parameters(c, circ, td)
    for r_, s_ in ((0, 0), (17, 0), (0, 23), (c.r - 1, c.r - 2)):
        A, B, C = cg.prove(pk, circ, r_, s_, [])
        exp = cg.trapdoor_expected(pk, circ, td, r_, s_, [])
        assert (A, B, C) == (exp["A"], exp["B"], exp["C"])
