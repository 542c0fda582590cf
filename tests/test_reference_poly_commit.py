"""The reference's own acceptance tests of the path, restated: cp-groth16/src/lib.rs `multi_stage_test::
poly_commit_test` (lib.rs:139-180) and `single_stage_test::poly_commit_test` (lib.rs:273-313) on the reference's
PolyEvalCircuit (oracle/reference_circuits.py): setup -> commit -> prove -> verify_proof(&pvk, &proof, &[point,
evaluation]).  The CPU half pins the oracle (generator + committer + prover + pairing verifier agree on the
reference's circuit); the GPU half runs the same flow through the C ABI and must produce the identical proof."""
import random

import pytest

from oracle import cpgroth16 as og
from oracle import pairing as pr
from oracle.params import get_curve
from oracle.reference_circuits import poly_eval_circuit


def _setup(name, stages, seed):
    c = get_curve(name)
    rng = random.Random(seed)
    degree = 10                                                        # lib.rs:144
    polynomial = [rng.randrange(c.r) for _ in range(degree)] + [1]     # random monic polynomial (lib.rs:145-146)
    point = rng.randrange(c.r)
    circ, inputs = poly_eval_circuit(name, polynomial, point, stages)
    td = og.sample_trapdoor(name, stages, seed + 1)
    pk = og.generate_parameters(name, circ, td)
    return c, circ, inputs, td, pk, rng


@pytest.mark.parametrize("name", ["bls12_381", "bn254"])
@pytest.mark.parametrize("stages", [2, 1])
def test_poly_commit_test_oracle(name, stages):
    c, circ, inputs, td, pk, rng = _setup(name, stages, 40 + stages)
    kappas = [rng.randrange(c.r)] if stages == 2 else []
    comms = [og.commit(pk, circ, 0, kappas[0])] if stages == 2 else []
    A, B, C = og.prove(pk, circ, rng.randrange(c.r), rng.randrange(c.r), kappas)
    pvk = pr.prepare_verifying_key(pk)
    assert pr.verify_proof(pvk, A, B, C, comms, inputs)                # lib.rs:178-179 / 311-312
    assert not pr.verify_proof(pvk, A, B, C, comms, [inputs[0], (inputs[1] + 1) % c.r])


@pytest.mark.gpu
@pytest.mark.parametrize("stages", [2, 1])
def test_poly_commit_test_gpu(ctx, stages):
    from hekaton_system_b200 import cp_groth16 as host
    from hekaton_system_b200 import rng as hrng
    cv = ctx.curve
    c, circ, inputs, td, opk, rng = _setup(cv.name, stages, 60 + stages)
    pk = host.ProvingKey.from_key(ctx, opk, circ)
    cb = host.CommitmentBuilder(host.StagedAssignment.from_circuit(circ, cv), pk)
    seeded = hrng.CurveRng(hrng.ChaChaRng(bytes([7] * 32), 20), cv)     # ark_std::test_rng() is a seeded ChaCha20
    comms, rands = [], []
    if stages == 2:
        com, rand = cb.commit(rng=seeded)                              # lib.rs:170
        comms, rands = [com], [rand]
    proof = cb.prove(comms, rands, rng=seeded)                         # lib.rs:174
    pvk = pr.prepare_verifying_key(opk)
    assert pr.verify_proof(pvk, proof.a, proof.b, proof.c, proof.ds, inputs)      # lib.rs:178-179
    # and it is exactly the oracle prover's proof for the same randomness
    ref = hrng.ChaChaRng(bytes([7] * 32), 20)
    kap = [ref.fr_rand(cv)] if stages == 2 else []
    r2, s2 = ref.fr_rand(cv), ref.fr_rand(cv)
    assert (proof.a, proof.b, proof.c) == og.prove(opk, circ, r2, s2, kap)
    if stages == 2:
        assert proof.ds == [og.commit(opk, circ, 0, kap[0])]
