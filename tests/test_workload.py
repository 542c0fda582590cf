"""The benchmark workload generator builds a VALID CP-Groth16 key: its key scalars equal the oracle's
restatement of generator.rs:18-238, and its expected proof exponents equal the oracle's trapdoor KAT."""
import pytest

from hekaton_system_b200 import workload
from oracle import poly, cpgroth16 as og
from oracle.params import get_curve


@pytest.mark.parametrize("cname", ["bls12_381", "bn254"])
def test_workload_matches_oracle_generator(cname):
    c = get_curve(cname)
    wl = workload.make_workload(cname, "tiny", seed=4, num_assignments=2)
    td, ks = workload.key_scalars(wl, seed=3)
    r1 = poly.R1CS(wl.l, wl.nw, *wl.rows)
    for z in wl.assignments:
        assert r1.is_satisfied(z, c.r)
    z = wl.assignments[1]
    circ = og.StagedCircuit(r1, list(zip(wl.stage_start[:-1], wl.stage_start[1:])), z[:wl.l], z[wl.l:])
    otd = og.Trapdoor(td["alpha"], td["beta"], td["gamma"], td["deltas"], td["tau"], td["g1"], td["g2"])
    opk = og.generate_parameters(c, circ, otd, with_points=False)
    assert opk.exps["a_g"] == ks["a_g"] and opk.exps["b_g"] == ks["b_g"]
    assert opk.exps["h_g"] == ks["h_g"] and opk.exps["deltas_abc_g"] == ks["ck"]
    assert opk.exps["gamma_abc_g"] == ks["gamma_abc"]
    exp = og.trapdoor_expected(opk, circ, otd, 11, 13, [17])["exps"]
    mine = workload.expected_exponents(wl, z, 11, 13, [17])
    assert (exp["A"], exp["B"], exp["C"], exp["D"]) == (mine["A"], mine["B"], mine["C"], mine["D"])


def test_shapes_have_the_documented_domain_sizes():
    for shape, log_n in (("big-merkle-small", 16), ("big-merkle", 18), ("vm", 21)):
        n, _ = workload.SHAPES[shape]
        assert (n + 4 - 1).bit_length() == log_n
