"""Oracle NTT / witness-map conventions: fast path == O(N^2) definition, round trips, coset."""
import random

import pytest

from oracle.params import BLS12_381, BN254
from oracle import poly, cpgroth16 as cg


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_fft_matches_definition(c):
    rng = random.Random(2)
    for n in (1, 2, 8, 64):
        d = poly.Domain(c, n)
        x = [rng.randrange(c.r) for _ in range(n)]
        assert d.fft(x) == d.dft_naive(x)
        assert d.ifft(d.fft(x)) == x
        co = d.get_coset(c.fr_generator)
        assert co.fft(x) == co.dft_naive(x)
        assert co.ifft(co.fft(x)) == x
        assert co.evaluate_vanishing_polynomial_on_coset() == d.evaluate_vanishing_polynomial(c.fr_generator)


@pytest.mark.parametrize("c", [BLS12_381, BN254], ids=lambda c: c.name)
def test_witness_map_identity(c):
    """h satisfies a(x) b(x) - c(x) = h(x) (x^N - 1) at a random point, and h[N-1] = 0."""
    circ = cg.synthetic_circuit(c, 50, stage0_witnesses=6, seed=9)
    z = circ.full_assignment()
    assert circ.r1cs.is_satisfied(z, c.r)
    dom, a, b, cc = poly.qap_evals(c, circ.r1cs, z)
    h = poly.h_from_evals(c, dom, a, b, cc)
    assert len(h) == dom.size and h[-1] == 0
    x = random.Random(4).randrange(c.r)
    ev = lambda co: sum(cf * pow(x, i, c.r) for i, cf in enumerate(co)) % c.r
    A, B, Cc = ev(dom.ifft(a)), ev(dom.ifft(b)), ev(dom.ifft(cc))
    assert (A * B - Cc) % c.r == ev(h) * (pow(x, dom.size, c.r) - 1) % c.r
