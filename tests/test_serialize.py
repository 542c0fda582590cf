"""Proof / point encodings (ark-serialize 0.4 as instantiated by ark-bn254 and ark-bls12-381 0.4): structure,
flags, round trips, and the published zcash-format encodings of the BLS12-381 generators as external vectors."""
from hekaton_system_b200 import cp_groth16 as host
from hekaton_system_b200.curves import BLS12_381, BN254
from oracle.ec import groups

# zcash / IETF "pairing-friendly curves" serialisation of the standard generators (public test vectors)
G1_GEN_COMPRESSED = ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                     "6c55e83ff97a1aeffb3af00adb22c6bb")
G1_GEN_UNCOMPRESSED = ("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                       "6c55e83ff97a1aeffb3af00adb22c6bb"
                       "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3ed"
                       "d03cc744a2888ae40caa232946c5e7e1")
G2_GEN_COMPRESSED = ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
                     "334cf11213945d57e5ac7d055d042b7e"
                     "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d177"
                     "0bac0326a805bbefd48056c8c121bdb8")


def test_bls12_381_zcash_generator_vectors():
    cv = BLS12_381
    G1, G2 = groups(cv.name)
    assert host.serialize_g1(G1.gen, cv, True).hex() == G1_GEN_COMPRESSED
    assert host.serialize_g1(G1.gen, cv, False).hex() == G1_GEN_UNCOMPRESSED
    assert host.serialize_g2(G2.gen, cv, True).hex() == G2_GEN_COMPRESSED
    unc = host.serialize_g2(G2.gen, cv, False)
    assert unc[:96].hex() == "1" + G2_GEN_COMPRESSED[1:] and len(unc) == 192       # same x, no flag bits
    # -G: same x, the sort flag (0x20) set when compressed; infinity: 0xc0 / 0x40 then zeros
    neg = host.serialize_g1(G1.neg(G1.gen), cv, True)
    assert neg[0] == 0x97 | 0x20 and neg[1:].hex() == G1_GEN_COMPRESSED[2:]
    assert host.serialize_g1(None, cv, True) == bytes([0xC0]) + bytes(47)
    assert host.serialize_g1(None, cv, False) == bytes([0x40]) + bytes(95)
    assert host.serialize_g2(None, cv, True) == bytes([0xC0]) + bytes(95)
    assert host.deserialize_g1_uncompressed(bytes.fromhex(G1_GEN_UNCOMPRESSED), cv) == G1.gen
    assert host.deserialize_g2_uncompressed(unc, cv) == G2.gen


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
        assert len(pr.serialize(cv, compress=True)) == nb + 2 * nb + nb + 8 + 2 * nb
        sa, sc = host.serialize_g1(A, cv, True), host.serialize_g1(C, cv, True)
        if cv.name == "bn254":
            # A and -A share x and differ exactly in the sign flag of the last byte
            assert sa[:-1] == sc[:-1] and (sa[-1] ^ sc[-1]) == 0x80
            assert host.serialize_g1(None, cv, True)[-1] == 0x40 and len(host.serialize_g1(None, cv, True)) == nb
            # field bytes are canonical little-endian: x < p after masking the flag bits
            assert int.from_bytes(sa[:-1] + bytes([sa[-1] & 0x3F]), "little") == A[0]
        else:
            assert sa[1:] == sc[1:] and (sa[0] ^ sc[0]) == 0x20 and sa[0] & 0x80
            assert int.from_bytes(bytes([sa[0] & 0x1F]) + sa[1:], "big") == A[0]
