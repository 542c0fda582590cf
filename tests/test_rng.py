"""Host RNG: the ChaCha core is pinned against RFC 7539; the rand_chacha / ark-ff conventions layered on it are
restated from their published source and are UNPINNED here (no Rust toolchain to produce vectors)."""
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
