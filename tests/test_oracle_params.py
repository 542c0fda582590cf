"""Pins the oracle's constants against public facts (generators on-curve, order r, roots of
unity) -- the reference itself holds no golden vectors for this path."""
import pytest

from oracle.params import BLS12_381, BN254
from oracle.ec import groups, ark_make_digits, ark_window_size, msm_pippenger_ark


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_generators_on_curve_and_order(c):
    G1, G2 = groups(c)
    assert G1.on_curve(G1.gen) and G2.on_curve(G2.gen)
    assert G1.mul(G1.gen, c.r) is None
    assert G2.mul(G2.gen, c.r) is None
    assert G1.mul(G1.gen, c.r - 1) == G1.neg(G1.gen)


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_two_adic_root(c):
    w = c.two_adic_root
    assert pow(w, 1 << c.two_adicity, c.r) == 1
    assert pow(w, 1 << (c.two_adicity - 1), c.r) == c.r - 1
    assert (c.r - 1) % (1 << c.two_adicity) == 0 and ((c.r - 1) >> c.two_adicity) % 2 == 1
    # multiplicative generator: g^((r-1)/2) = -1
    assert pow(c.fr_generator, (c.r - 1) // 2, c.r) == c.r - 1


def test_known_constants():
    # SURVEY.md 8c (values of ark-bls12-381 / ark-bn254 0.4)
    assert BLS12_381.two_adic_root == 10238227357739495823651030575849232062558860180284477541189508159991286009131
    assert BN254.two_adic_root == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert (-pow(BLS12_381.p, -1, 1 << 64)) % (1 << 64) == 0x89f3fffcfffcfffd
    assert (-pow(BLS12_381.r, -1, 1 << 32)) % (1 << 32) == 0xffffffff
    assert (-pow(BN254.p, -1, 1 << 32)) % (1 << 32) == 0xe4866389
    assert (-pow(BN254.r, -1, 1 << 32)) % (1 << 32) == 0xefffffff
    # Montgomery encodings as published in zkcrypto/bls12_381 `Scalar` (GENERATOR, ROOT_OF_UNITY)
    R = 1 << 256
    assert 7 * R % BLS12_381.r == 0x351332208fc5a8c4ff9c57876f8457b017e363d300189c0f0000000efffffff1
    assert BLS12_381.two_adic_root * R % BLS12_381.r == \
        0x5bf3adda19e9b27b0af53ae352a31e645b1b4c801819d7ecb9b58d8c5f0e466a


def test_ark_digits_and_window():
    import random
    rng = random.Random(5)
    assert ark_window_size(31) == 3 and ark_window_size(1 << 20) == 15
    for w in (3, 7, 15, 16):
        for _ in range(50):
            a = rng.randrange(BLS12_381.r)
            d = ark_make_digits(a, w, 255)
            assert sum(x << (i * w) for i, x in enumerate(d)) == a
            assert all(-(1 << (w - 1)) <= x < (1 << (w - 1)) for x in d[:-1])


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_pippenger_matches_naive(c):
    import random
    rng = random.Random(11)
    G1, G2 = groups(c)
    for G, n in ((G1, 40), (G2, 8)):
        bases = [G.mul(G.gen, rng.randrange(1, c.r)) for _ in range(n)]
        bases[3] = None
        scalars = [rng.randrange(c.r) for _ in range(n)]
        scalars[5] = 0
        scalars[6] = 1
        scalars[7] = c.r - 1
        assert msm_pippenger_ark(G, bases, scalars, c.fr_bits) == G.msm_naive(bases, scalars)
