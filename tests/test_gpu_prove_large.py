"""GPU parity at BASELINE.json's full size (big-merkle shape, n = 200 000, N = 2^18): the proof of
the device path equals the trapdoor known answer (exact expected A, B, C, D from O(m) scalar
arithmetic and one scalar multiplication each -- no MSM or NTT on the checking side), and the
commitment/proof pass the reference's structural checks."""
import pytest

from hekaton_system_b200 import codec, workload
from hekaton_system_b200.cp_groth16 import CommitmentBuilder
from oracle.ec import groups

pytestmark = pytest.mark.gpu


class _Assign:
    def __init__(self, z):
        self.z = z

    def full_assignment(self):
        return self.z


@pytest.mark.parametrize("shape", ["big-merkle-small", "big-merkle"])
def test_full_size_proof_matches_trapdoor(ctx, shape):
    cv = ctx.curve
    if cv.name != "bls12_381" and shape == "big-merkle":
        pytest.skip("BN254 at full size: tests/test_gpu_configs.py::test_big_merkle_full_size_bn254_batch")
    wl = workload.make_workload(cv, shape, seed=2, num_assignments=1)
    td, _ = workload.key_scalars(wl, seed=5)
    pk = workload.build_device_key(ctx, wl)
    z = wl.assignments[0]
    kappa, r_, s_ = 0x123456789abcdef0123456789, cv.r - 12345, 0xfeedface
    cb = CommitmentBuilder(_Assign(z), pk)
    com, _ = cb.commit(kappa=kappa)
    proof = cb.prove([com], [kappa], r=r_, s=s_)
    exp = workload.expected_exponents(wl, z, r_, s_, [kappa])
    G1, G2 = groups(cv.name)
    G = G1.mul(G1.gen, td["g1"])
    H = G2.mul(G2.gen, td["g2"])
    assert com == G1.mul(G, exp["D"][0])
    assert proof.a == G1.mul(G, exp["A"])
    assert proof.b == G2.mul(H, exp["B"])
    assert proof.c == G1.mul(G, exp["C"])
    if shape == "big-merkle-small":
        # independent acceptance check at N = 2^16: verify_proof (verifier.rs:23-71) through the pairing oracle
        from oracle import pairing as pr
        from types import SimpleNamespace
        ks = wl.key_scalars
        vk = SimpleNamespace(curve=cv.name, alpha_g=G1.mul(G, ks["alpha"]), beta_h=G2.mul(H, ks["beta"]),
                             gamma_h=G2.mul(H, td["gamma"]), deltas_h=[G2.mul(H, d) for d in td["deltas"]],
                             gamma_abc_g=[G1.mul(G, x) for x in ks["gamma_abc"]])
        assert pr.verify_proof(pr.prepare_verifying_key(vk), proof.a, proof.b, proof.c, [com], z[1:wl.l])
    # h[N-1] == 0 at full size (prover.rs:128-129) through the resident-matrix witness map
    import numpy as np
    N = 1 << wl.log_n
    # re-upload matrices for the standalone call
    r1 = ctx.r1cs_upload(wl.n, wl.l, wl.nw, wl.csr)
    h = ctx.witness_map(r1, codec.to_mont(z, cv.r, 4), n_out=N)
    assert not h[N - 1].any() and h[:N - 1].any()
    ctx.r1cs_free(r1)
