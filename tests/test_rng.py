"""Host RNG.  Pinned: the ChaCha20 core against RFC 7539 2.3.2; the 8- / 12- / 20-round variants against the
ECRYPT zero-key keystream vectors; the word order, block order and counter position of the buffered stream against
rand_chacha's own `test_chacha_true_values_a` (seed 0: the first two blocks as 32 u32 words = RFC 7539 A.1 vectors
1 and 2).  Still unpinned (no Rust toolchain to produce vectors): ark-ff's `Fr::rand` shave / reject rule."""
from hekaton_system_b200 import rng
from hekaton_system_b200.curves import BLS12_381, BN254


def test_chacha20_block_rfc7539_2_3_2():
    key = list(range(32))
    import struct
    kw = struct.unpack("<8I", bytes(key))
    # counter = 1, nonce = 00:00:00:09:00:00:00:4a:00:00:00:00
    out = rng.chacha_block(kw, [1, 0x09000000, 0x4A000000, 0], 20)
    expect = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
              0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert out == expect


def test_round_variants_ecrypt_zero_key_vectors():
    exp = {
        20: "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
            "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586",
        12: "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
            "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be",
        8: "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
           "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42",
    }
    for rounds, hexs in exp.items():
        assert rng.ChaChaRng(bytes(32), rounds).fill_bytes(64).hex() == hexs


def test_rand_chacha_true_values_a():
    """rand_chacha 0.3 `test_chacha_true_values_a`: ChaCha20Rng::from_seed([0; 32]), 16 next_u32 twice."""
    g = rng.ChaChaRng(bytes(32), 20)
    first = [0xade0b876, 0x903df1a0, 0xe56a5d40, 0x28bd8653, 0xb819d2bd, 0x1aed8da0, 0xccef36a8, 0xc70d778b,
             0x7c5941da, 0x8d485751, 0x3fe02477, 0x374ad8b8, 0xf4b8436a, 0x1ca11815, 0x69b687c3, 0x8665eeb2]
    second = [0xbee7079f, 0x7a385155, 0x7c97ba98, 0x0d082d73, 0xa0290fcb, 0x6965e348, 0x3e53c612, 0xed7aee32,
              0x7621b729, 0x434ee69c, 0xb03371d5, 0xd539d874, 0x281fed31, 0x45fb0a51, 0x1f0ae1ac, 0x6f4d794b]
    assert [g.next_u32() for _ in range(16)] == first
    assert [g.next_u32() for _ in range(16)] == second


def test_stream_structure_and_fr_rand_range():
    r1 = rng.ChaChaRng(bytes([69] * 32), 12)       # the reference's setup seed (data_structures.rs:54-55)
    r2 = rng.ChaChaRng(bytes([69] * 32), 12)
    a = [r1.next_u64() for _ in range(40)]
    b = [r2.next_u32() for _ in range(80)]
    assert a == [b[2 * i] | (b[2 * i + 1] << 32) for i in range(40)]      # next_u64 = two LE words, across refills
    assert rng.ChaChaRng(bytes([69] * 32), 20).next_u64() != a[0]
    for cv in (BLS12_381, BN254):
        g = rng.ChaChaRng(bytes(range(32)), 12)
        for _ in range(200):
            m = g.fr_rand_mont(cv)
            assert 0 <= m < cv.r and m < (1 << cv.fr_bits)
        g1, g2 = rng.ChaChaRng(bytes(range(32)), 12), rng.ChaChaRng(bytes(range(32)), 12)
        assert g1.fr_rand(cv) * cv.fr_R % cv.r == g2.fr_rand_mont(cv)


def test_commitment_builder_draw_order():
    """commit draws kappa; prove draws r then s (committer.rs:85, prover.rs:28-29)."""
    cv = BN254
    base = rng.ChaChaRng(bytes([7] * 32), 12)
    exp = [base.fr_rand(cv) for _ in range(3)]
    crng = rng.CurveRng(rng.ChaChaRng(bytes([7] * 32), 12), cv)
    assert [crng.fr_rand(), crng.fr_rand(), crng.fr_rand()] == exp
