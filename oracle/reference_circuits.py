"""The circuits of the reference's own cp-groth16 tests, as R1CS (oracle; TEST INFRASTRUCTURE ONLY).

/root/reference/cp-groth16/src/lib.rs holds the two acceptance tests of the path:
  multi_stage_test::poly_commit_test   (lib.rs:40-181)   PolyEvalCircuit with 2 stages
  single_stage_test::poly_commit_test  (lib.rs:183-313)  the same statement in 1 stage
Both: sample a monic polynomial of degree 10, setup, (commit,) prove, `verify_proof(&pvk, &proof, &[point,
evaluation])`.  The circuit (lib.rs:81-115 / 226-252):
  stage 0   witness the coefficients c_0..c_d; enforce c_d == 1 (twice in the 2-stage version, lib.rs:87-94)
  stage 1   inputs `point`, `evaluation`; claimed = sum_i c_i * point^i with `coeff * &cur_pow` and
            `cur_pow *= &point` per coefficient; enforce claimed == evaluation
The gadget layer (ark-r1cs-std FpVar) is not available here, so the constraint system is written out directly with
the semantics FpVar has: constant * variable and sums are linear combinations (no constraint), variable * variable
allocates a product witness with one constraint, `enforce_equal(a, b)` is (a - b) * 1 = 0.  Variable numbering is
ours (stage-0 witnesses first, then stage-1 witnesses in allocation order), which is all the prover's arithmetic
depends on; the test then reads like the reference's.
"""
from . import poly
from .cpgroth16 import StagedCircuit
from .params import get_curve


def poly_eval_circuit(curve, polynomial, point, stages=2):
    """-> (StagedCircuit, public_inputs = [point, evaluation]).  `polynomial`: coefficients, low degree first, monic."""
    c = get_curve(curve)
    r = c.r
    assert stages in (1, 2) and polynomial[-1] % r == 1
    d = len(polynomial) - 1
    evaluation = sum(cf * pow(point, i, r) for i, cf in enumerate(polynomial)) % r
    l = 3                                            # one, point, evaluation
    ONE, POINT, EVAL = 0, 1, 2
    coeff = [l + i for i in range(d + 1)]            # stage-0 witnesses
    witness = [cf % r for cf in polynomial]
    a, b, cc = [], [], []

    def enforce_equal(lc1, lc2):                     # (lc1 - lc2) * 1 = 0
        a.append(lc1 + [((r - cf) % r, j) for cf, j in lc2])
        b.append([(1, ONE)])
        cc.append([])

    enforce_equal([(1, coeff[d])], [(1, ONE)])
    if stages == 2:
        enforce_equal([(1, coeff[d])], [(1, ONE)])   # the 2-stage test enforces it twice (lib.rs:87-94)
    n_stage0 = len(witness)

    def new_witness(value):
        witness.append(value % r)
        return l + len(witness) - 1

    def mul(lc_x, val_x, lc_y, val_y):               # variable * variable -> fresh witness, one constraint
        w = new_witness(val_x * val_y)
        a.append(lc_x)
        b.append(lc_y)
        cc.append([(1, w)])
        return [(1, w)], val_x * val_y % r

    claimed = []                                     # linear combination
    cur_pow, cur_val, cur_const = [(1, ONE)], 1, True
    for i in range(d + 1):
        if cur_const:                                # coeff * constant one: a linear combination
            claimed += [(1, coeff[i])]
        else:
            term, _ = mul([(1, coeff[i])], witness[i], cur_pow, cur_val)
            claimed += term
        if cur_const:                                # one * point: still linear
            cur_pow, cur_val, cur_const = [(1, POINT)], point % r, False
        else:
            cur_pow, cur_val = mul(cur_pow, cur_val, [(1, POINT)], point % r)
    enforce_equal(claimed, [(1, EVAL)])
    r1cs = poly.R1CS(l, len(witness), a, b, cc)
    instance = [1, point % r, evaluation]
    ranges = [(0, n_stage0), (n_stage0, len(witness))] if stages == 2 else [(0, len(witness))]
    circ = StagedCircuit(r1cs, ranges, instance, witness)
    assert r1cs.is_satisfied(circ.full_assignment(), r)
    return circ, [point % r, evaluation]
