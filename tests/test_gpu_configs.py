"""GPU parity at the sizes BASELINE.json's configs name (driver-run tier):

  config 2   standalone Fr NTT 2^16 / 2^20 / 2^21 / 2^24, all four modes, compared ELEMENT-WISE with the C
             oracle's Radix2EvaluationDomain restatement (oracle/c/oracle.c: fr_ntt); G1 MSM 2^22 / 2^24 and
             G2 MSM 2^20 / 2^22 checked exactly (bases k_i * G, expected point = (sum k_i s_i mod r) * G)
  config 4   VKD shape (N = 2^14), a batch of 64 subcircuit proofs vs the trapdoor known answer
  config 5   VM shape (N = 2^21, 2.2e5-point stage-0 commitment), a batch of 4 proofs vs the trapdoor
  config 3   big-merkle full size on BN254 as well (the curve the reference's binaries instantiate)
"""
import random

import numpy as np
import pytest

from hekaton_system_b200 import codec, workload
from oracle import cref
from oracle.ec import groups

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("log_n", [12, 14, 15, 16, 17, 18, 19, 20, 21, 24])
def test_ntt_elementwise_vs_c_oracle(ctx, log_n):
    """Every output element of hk_ntt equals ark-poly's fft / ifft / coset fft / coset ifft as restated in C
    (in-order radix-2 with the bit reversal up front -- a different schedule from the device's tiled radix-8
    passes).  2^21 and 2^24 exercise the 3-pass plans and the top-passenger tile; 2^12 ... 2^18 the smaller tiles a
    single transform gets when 2^11-element tiles could not fill the GPU (ntt_tile_bits)."""
    cv = ctx.curve
    n = 1 << log_n
    rng = np.random.default_rng(100 + log_n)
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    x[:, 3] >>= 2                                   # < 2^252 < r on both curves (raw Montgomery residues)
    x[0] = 0
    x[1] = codec.to_mont([1], cv.r, 4)[0]
    x[2] = codec.to_mont([cv.r - 1], cv.r, 4)[0]
    for inverse, coset in ((False, False), (True, False), (False, True), (True, True)):
        got = x.copy()
        ctx.ntt(got, log_n, inverse=inverse, coset=coset)
        ref = x.copy()
        cref.ntt(cv.cid, ref, log_n, inverse=inverse, coset=coset, threads=cref.max_threads())
        assert np.array_equal(got, ref), (log_n, inverse, coset)


@pytest.mark.parametrize("log_n,K", [(13, 5), (16, 3), (16, 40)])
def test_ntt_batch_elementwise_vs_c_oracle(ctx, log_n, K):
    """K transforms in one call: small batches take the small tiles, large ones the 2^11-element tiles."""
    cv = ctx.curve
    n = 1 << log_n
    rng = np.random.default_rng(7 * log_n + K)
    x = rng.integers(0, 1 << 62, size=(K * n, 4), dtype=np.uint64)
    x[:, 3] >>= 2
    for inverse, coset in ((False, True), (True, False)):
        got = x.copy()
        ctx.ntt(got, log_n, inverse=inverse, coset=coset, K=K)
        for k in range(K):
            ref = x[k * n:(k + 1) * n].copy()
            cref.ntt(cv.cid, ref, log_n, inverse=inverse, coset=coset, threads=cref.max_threads())
            assert np.array_equal(got[k * n:(k + 1) * n], ref), (log_n, K, k, inverse, coset)


@pytest.mark.parametrize("log_n", [16, 18])
def test_witness_map_elementwise_vs_c_oracle(ctx, log_n):
    """hk_qap_witness_map (resident CSR matrices -> 7 fused transforms) == the C restatement of
    LibsnarkReduction::witness_map_from_matrices on every coefficient, at the big-merkle domain sizes."""
    cv = ctx.curve
    shape = "big-merkle-small" if log_n == 16 else "big-merkle"
    wl = workload.make_workload(cv, shape, seed=3, num_assignments=1)
    assert wl.log_n == log_n
    z = codec.to_mont(wl.assignments[0], cv.r, 4)
    r1 = ctx.r1cs_upload(wl.n, wl.l, wl.nw, wl.csr)
    h = ctx.witness_map(r1, z, n_out=1 << log_n)
    ref = cref.witness_map(cv.cid, wl.n, wl.l, wl.csr, z, threads=cref.max_threads())
    assert np.array_equal(h, ref)
    ctx.r1cs_free(r1)


# ---------------------------------------------------------------------------------------- MSM
def _rand_scalars(rng, n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] >>= 2                                   # < 2^253 < r on both curves
    return a


@pytest.mark.parametrize("group,log_n,kind", [(1, 22, "uniform"), (1, 24, "uniform"), (1, 24, "witness"),
                                              (2, 20, "uniform"), (2, 22, "uniform")])
def test_msm_sweep_sizes_exact(ctx, group, log_n, kind):
    """BASELINE config 2's upper sizes.  Expected exponent from the C oracle's Fr dot product (no MSM on the
    checking side), expected point from one fixed-base multiplication in the Python oracle."""
    cv = ctx.curve
    if cv.name != "bls12_381" and log_n > 22 - 2 * (group - 1):
        pytest.skip("largest sizes on the primary curve; BN254 runs G1 2^22 / G2 2^20")
    G = groups(cv.name)[group - 1]
    n = 1 << log_n
    rng = np.random.default_rng(1000 * group + log_n)
    ks = _rand_scalars(rng, n)
    enc = codec.g1_affine_to_limbs if group == 1 else codec.g2_affine_to_limbs
    dec = codec.g1_affine_from_limbs if group == 1 else codec.g2_affine_from_limbs
    h = ctx.bases_from_scalars(group, enc([G.gen], cv)[0], ks)
    sc = _rand_scalars(rng, n)
    if kind == "witness":                            # 45 % zeros, 45 % ones, 10 % uniform
        u = rng.random(n)
        sc[u >= 0.1] = 0
        sc[(u >= 0.1) & (u < 0.55), 0] = 1
    sc[0] = codec.ints_to_limbs([cv.r - 1], 4)[0]
    sc[1] = 0
    e = cref.dot_fr(cv.cid, ks, sc)
    exp = G.mul(G.gen, e)
    for table in (False, True):
        if table:
            ctx.bases_precompute(h, 0.1 if kind == "witness" else 1.0)
        aff, inf = ctx.into_affine(group, ctx.msm(group, h, sc))
        assert dec(aff, cv, inf)[0] == exp, (group, log_n, kind, table)
    ctx.bases_free(h)


# ---------------------------------------------------------------------------------------- proofs
def _check_batch(ctx, wl, pk, K, seed):
    """K proofs (hk_commit_batch + hk_prove_batch) against the trapdoor known answer."""
    cv = ctx.curve
    td = wl.trapdoor
    G1, G2 = groups(cv.name)
    G = G1.mul(G1.gen, td["g1"])
    H = G2.mul(G2.gen, td["g2"])
    rng = random.Random(seed)
    zs_m = [codec.to_mont(z, cv.r, 4) for z in wl.assignments]
    idx = [k % len(zs_m) for k in range(K)]
    rs = [rng.randrange(cv.r) for _ in range(K)]
    ss = [rng.randrange(cv.r) for _ in range(K)]
    ks = [rng.randrange(cv.r) for _ in range(K)]
    rs[0], ss[min(1, K - 1)] = 0, 0                   # prove_last_stage_without_zk / the r == 0 branch (prover.rs:92)
    fr = lambda v: codec.to_mont(v, cv.r, 4)
    z = np.stack([zs_m[i] for i in idx])
    com, cinf = ctx.commit_batch(pk.handle, 0, z, fr(ks))
    a, b, c, inf = ctx.prove_batch(pk.handle, z, fr(rs), fr(ss), fr(ks).reshape(K, 1, 4))
    assert not cinf.any() and not inf.any()
    # the expected exponents are linear in (r, s, kappa): evaluate the O(m) part once per distinct assignment
    cache = {}
    for k in range(K):
        exp = workload.expected_exponents(wl, wl.assignments[idx[k]], rs[k], ss[k], [ks[k]], cache=cache.setdefault(idx[k], {}))
        assert codec.g1_affine_from_limbs(com[k], cv)[0] == G1.mul(G, exp["D"][0]), k
        assert codec.g1_affine_from_limbs(a[k], cv)[0] == G1.mul(G, exp["A"]), k
        assert codec.g2_affine_from_limbs(b[k], cv)[0] == G2.mul(H, exp["B"]), k
        assert codec.g1_affine_from_limbs(c[k], cv)[0] == G1.mul(G, exp["C"]), k


def test_vkd_shape_batch_of_64(ctx):
    """BASELINE config 4 (VKD, N = 2^14): the small-N regime runs as one 64-proof pipeline."""
    wl = workload.make_workload(ctx.curve, "vkd", seed=11, num_assignments=4)
    workload.key_scalars(wl, seed=12)
    pk = workload.build_device_key(ctx, wl)
    _check_batch(ctx, wl, pk, 64, 13)
    ctx.pk_free(pk.handle)


def test_vm_shape_batch_of_4(ctx):
    """BASELINE config 5 (VM, 1024 cycles per subcircuit: n = 1.14 M, N = 2^21, stage-0 commitment over
    2.2e5 points)."""
    cv = ctx.curve
    if cv.name != "bls12_381":
        pytest.skip("VM full size on the primary curve (host-side workload generation takes ~1 min per curve)")
    wl = workload.make_workload(cv, "vm", seed=21, num_assignments=2)
    assert wl.log_n == 21 and wl.stage_start[1] > 200_000
    workload.key_scalars(wl, seed=22)
    pk = workload.build_device_key(ctx, wl)
    _check_batch(ctx, wl, pk, 4, 23)
    ctx.pk_free(pk.handle)


def test_big_merkle_full_size_bn254_batch(ctx):
    """BASELINE config 3's subcircuit shape on the curve the reference's binaries use."""
    cv = ctx.curve
    if cv.name != "bn254":
        pytest.skip("BLS12-381 full size: tests/test_gpu_prove_large.py")
    wl = workload.make_workload(cv, "big-merkle", seed=31, num_assignments=2)
    workload.key_scalars(wl, seed=32)
    pk = workload.build_device_key(ctx, wl)
    _check_batch(ctx, wl, pk, 8, 33)
    ctx.pk_free(pk.handle)
