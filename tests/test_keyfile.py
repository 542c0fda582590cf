"""Key-file reader / writer: ProvingKeys::serialize_uncompressed layout
(/root/reference/mpi-snark/src/data_structures.rs:41-50, cp-groth16/src/data_structures.rs:32-114).
Unpinned against the reference (it ships no key file): round trips, layout arithmetic, flag rules."""
import struct

import numpy as np
import pytest

from hekaton_system_b200 import codec, keyfile
from hekaton_system_b200.curves import BLS12_381, BN254
from oracle import cpgroth16 as og
from oracle.params import BLS12_381 as O_BLS, BN254 as O_BN


def _small_key(oc, seed):
    circ = og.synthetic_circuit(oc, 40, stage0_witnesses=5, seed=seed)
    td = og.sample_trapdoor(oc, 2, seed + 1)
    return circ, td, og.generate_parameters(oc, circ, td)


@pytest.mark.parametrize("cv,oc", [(BLS12_381, O_BLS), (BN254, O_BN)], ids=["bls12_381", "bn254"])
def test_roundtrip_and_layout(cv, oc):
    circ, td, pk = _small_key(oc, 31)
    _, _, pk2 = _small_key(oc, 41)
    nb = 8 * cv.fq_limbs
    params = bytes(range(17))
    raw = keyfile.dumps_proving_keys("BigMerkle circuit", params, {0: pk, 3: pk2}, {0: 0, 1: 0, 2: 3, 3: 3}, cv)
    # header: String and Vec<u8> are u64 LE length + bytes, then the map length
    assert raw[:8] == struct.pack("<Q", 17) and raw[8:25] == b"BigMerkle circuit"
    assert raw[25:33] == struct.pack("<Q", 17) and raw[33:50] == params
    assert raw[50:58] == struct.pack("<Q", 2) and raw[58:66] == struct.pack("<Q", 0)

    def pk_size(k):
        g1, g2 = 2 * nb, 4 * nb
        n = g1 + 3 * g2 + 8 + len(k.gamma_abc_g) * g1 + 8 + len(k.deltas_h) * g2          # vk
        n += g1 + 8 + len(k.a_g) * g1 + 8 + len(k.b_g) * g1 + 8 + len(k.b_h) * g2 + 8 + len(k.h_g) * g1
        n += g1 + 8 + sum(8 + len(v) * g1 for v in k.deltas_abc_g) + 8 + len(k.deltas_g) * g1
        return n

    assert len(raw) == 50 + 8 + (8 + pk_size(pk)) + (8 + pk_size(pk2)) + 8 + 4 * 16
    back = keyfile.read_proving_keys(raw, cv)
    assert back.circuit_id == "BigMerkle circuit" and back.serialized_circ_params == params
    assert back.subcircuit_representative_map == {0: 0, 1: 0, 2: 3, 3: 3} and back.num_subcircuits() == 4
    assert sorted(back.minimal_proving_keys) == [0, 3]
    for k, ref in ((0, pk), (1, pk), (2, pk2), (3, pk2)):
        got = back.get_pk(k)
        assert got.a_g.to_points(1) == ref.a_g and got.b_g.to_points(1) == ref.b_g
        assert got.b_h.to_points(2) == ref.b_h and got.h_g.to_points(1) == ref.h_g
        assert [v.to_points(1) for v in got.deltas_abc_g] == ref.deltas_abc_g
        assert got.deltas_g.to_points(1) == ref.deltas_g and got.last_delta_g.to_points(1) == [ref.deltas_g[-1]]
        assert got.beta_g.to_points(1) == [ref.beta_g] and got.vk.alpha_g.to_points(1) == [ref.alpha_g]
        assert got.vk.beta_h.to_points(2) == [ref.beta_h] and got.vk.gamma_h.to_points(2) == [ref.gamma_h]
        assert got.vk.last_delta_h.to_points(2) == [ref.last_delta_h]
        assert got.vk.gamma_abc_g.to_points(1) == ref.gamma_abc_g and got.vk.deltas_h.to_points(2) == ref.deltas_h
    with pytest.raises(KeyError):
        back.get_pk(9)
    # b_g of a Groth16 key has points at infinity (variables absent from B): the mask must carry them
    assert any(p is None for p in pk.b_g) == bool(back.get_pk(0).b_g.inf.any())
    # a single-key blob, truncation and trailing bytes
    import io
    b = io.BytesIO()
    keyfile.write_proving_key(b, pk, cv)
    one = keyfile.read_proving_key(b.getvalue(), cv)
    assert one.a_g.to_points(1) == pk.a_g
    with pytest.raises(ValueError):
        keyfile.read_proving_key(b.getvalue()[:-1], cv)
    with pytest.raises(ValueError):
        keyfile.read_proving_key(b.getvalue() + b"\0", cv)


def test_infinity_flag_with_coordinates_is_rejected():
    cv = BN254
    nb = 8 * cv.fq_limbs
    bad = bytearray(2 * nb)
    bad[0] = 1
    bad[-1] |= 0x40
    r = keyfile._Reader(bytes(bad), cv)
    with pytest.raises(ValueError):
        r.point(1)


@pytest.mark.gpu
def test_prove_from_key_file(ctx, tmp_path):
    """A key written in the reference's file format, loaded with the bulk reader, uploaded (Montgomery
    conversion on the device) and used for commit + prove: the proof equals the trapdoor known answer."""
    from hekaton_system_b200 import cp_groth16 as host
    from oracle.params import get_curve
    oc = get_curve(ctx.curve.name)
    circ, td, opk = _small_key(oc, 51)
    path = tmp_path / "keys.bin"
    with open(path, "wb") as f:
        keyfile.write_proving_keys(f, "R1CS circuit", b"", {0: opk}, {0: 0, 1: 0}, ctx.curve)
    keys = keyfile.load(path, ctx.curve)
    staged = host.StagedAssignment.from_circuit(circ, ctx.curve)
    pk = keyfile.to_device(ctx, keys.get_pk(1), staged)
    cb = host.CommitmentBuilder(staged, pk)
    kappa, r_, s_ = 0x1234567, 0xabcdef0123, 0x77665544332211
    com, _ = cb.commit(kappa=kappa)
    proof = cb.prove([com], [kappa], r=r_, s=s_)
    exp = og.trapdoor_expected(opk, circ, td, r_, s_, [kappa])
    assert com == exp["D"][0]
    assert (proof.a, proof.b, proof.c) == (exp["A"], exp["B"], exp["C"])
