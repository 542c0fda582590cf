"""ark-serialize proof encoding (restated; unpinned -- see the module docstring): structure, flags, round trip."""
from hekaton_system_b200 import cp_groth16 as host
from hekaton_system_b200.curves import BLS12_381, BN254
from oracle.ec import groups


def test_proof_roundtrip_and_flags():
    for cv in (BLS12_381, BN254):
        G1, G2 = groups(cv.name)
        nb = 8 * cv.fq_limbs
        A, B, C, D = G1.mul(G1.gen, 5), G2.mul(G2.gen, 7), G1.neg(G1.mul(G1.gen, 5)), None
        pr = host.Proof(A, B, C, [D, A])
        raw = pr.serialize(cv)
        assert len(raw) == 2 * nb + 4 * nb + 2 * nb + 8 + 2 * (2 * nb)
        back = host.Proof.deserialize_uncompressed(raw, cv)
        assert (back.a, back.b, back.c, back.ds) == (A, B, C, [D, A])
        # A and -A share x and differ exactly in the sign flag of the last byte
        sa, sc = host.serialize_g1(A, cv, True), host.serialize_g1(C, cv, True)
        assert sa[:-1] == sc[:-1] and (sa[-1] ^ sc[-1]) == 0x80
        assert host.serialize_g1(None, cv, True)[-1] == 0x40 and len(host.serialize_g1(None, cv, True)) == nb
        assert len(pr.serialize(cv, compress=True)) == nb + 2 * nb + nb + 8 + 2 * nb
        # field bytes are canonical little-endian: x < p after masking the flag bits
        x = int.from_bytes(sa[:-1] + bytes([sa[-1] & 0x3F]), "little")
        assert x == A[0]
