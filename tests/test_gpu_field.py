"""GPU parity: element-wise field arithmetic through the C ABI vs Python big integers (bit-exact)."""
import random

import numpy as np
import pytest

from hekaton_system_b200 import codec
from hekaton_system_b200 import backend as B

pytestmark = pytest.mark.gpu


def _cases(mod, nlimbs, rng, n=600):
    R = 1 << (64 * nlimbs)
    edge = [0, 1, 2, mod - 1, mod - 2, R % mod, R * R % mod, (mod - 1) // 2, (1 << 64) - 1, mod - (1 << 64),
            (1 << (mod.bit_length() - 1)), (1 << 32) - 1]
    a = [e % mod for e in edge for _ in edge] + [rng.randrange(mod) for _ in range(n)]
    b = [e % mod for _ in edge for e in edge] + [rng.randrange(mod) for _ in range(n)]
    return a, b


@pytest.mark.parametrize("field", [B.FIELD_FR, B.FIELD_FQ])
def test_prime_field_ops(ctx, field):
    cv = ctx.curve
    mod, L = (cv.r, 4) if field == B.FIELD_FR else (cv.p, cv.fq_limbs)
    R = 1 << (64 * L)
    Rinv = pow(R, -1, mod)
    rng = random.Random(100 + field)
    a, b = _cases(mod, L, rng)
    A, Bm = codec.ints_to_limbs(a, L), codec.ints_to_limbs(b, L)
    # raw residues: mul computes a*b/R, add/sub are plain
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_MUL, A, Bm))
    assert got == [x * y * Rinv % mod for x, y in zip(a, b)]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_ADD, A, Bm))
    assert got == [(x + y) % mod for x, y in zip(a, b)]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_SUB, A, Bm))
    assert got == [(x - y) % mod for x, y in zip(a, b)]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_TO_MONT, A))
    assert got == [x * R % mod for x in a]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_FROM_MONT, A))
    assert got == [x * Rinv % mod for x in a]
    # inverse in Montgomery form: inv(xR) = x^-1 R ; 0 -> 0
    am = [x * R % mod for x in a[:200]]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_INV, codec.ints_to_limbs(am, L)))
    assert got == [(pow(x, -1, mod) * R % mod) if x % mod else 0 for x in a[:200]]
    # dedicated squaring (generated separately from the multiply): a^2 R^-1, random and top-of-range values
    sq = a[:400] + [mod - 1 - k for k in range(40)] + [(mod - 1) // 2, 1 << (mod.bit_length() - 1)]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_SQR, codec.ints_to_limbs(sq, L)))
    assert got == [x * x * pow(R, -1, mod) % mod for x in sq]
    # binary-GCD inverse (ALU-pipe variant used by the affine normalisations): same values
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_INV_GCD, codec.ints_to_limbs(am, L)))
    assert got == [(pow(x, -1, mod) * R % mod) if x % mod else 0 for x in a[:200]]
    # Karatsuba-product multiply (half-width edge values: equal / zero / all-ones halves) and the fused dual product
    # a*b - b*(a+b) (one reduction; operand p - b reaches p itself for b = 0)
    half = 1 << (32 * L)
    ke = [half - 1, half, (half - 1) * half % mod, half + 1, (half + 1) * (half - 1) % mod, mod - half]
    ka = a + [x % mod for x in ke for _ in ke]
    kb = b + [y % mod for _ in ke for y in ke]
    KA, KB = codec.ints_to_limbs(ka, L), codec.ints_to_limbs(kb, L)
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_MUL_KARA, KA, KB))
    assert got == [x * y * Rinv % mod for x, y in zip(ka, kb)]
    got = codec.limbs_to_ints(ctx.field_op(field, B.FOP_MULSUB2, KA, KB))
    assert got == [(x * y - y * ((x + y) % mod)) * Rinv % mod for x, y in zip(ka, kb)]


def test_fq2_ops(ctx):
    cv = ctx.curve
    p, L = cv.p, cv.fq_limbs
    R = 1 << (64 * L)
    rng = random.Random(7)
    n = 300
    a = [(rng.randrange(p), rng.randrange(p)) for _ in range(n)] + [(0, 0), (1, 0), (0, 1), (p - 1, p - 1)]
    b = [(rng.randrange(p), rng.randrange(p)) for _ in range(n)] + [(5, 7), (0, 0), (p - 1, 1), (p - 1, p - 1)]
    enc = lambda v: codec.to_mont([c for x in v for c in x], p, L).reshape(len(v), 2 * L)
    dec = lambda arr: [tuple(t) for t in np.array(codec.from_mont(arr.reshape(-1, L), p), dtype=object).reshape(-1, 2)]
    A, Bm = enc(a), enc(b)
    mul = lambda x, y: ((x[0] * y[0] - x[1] * y[1]) % p, (x[0] * y[1] + x[1] * y[0]) % p)
    assert dec(ctx.field_op(B.FIELD_FQ2, B.FOP_MUL, A, Bm)) == [mul(x, y) for x, y in zip(a, b)]
    assert dec(ctx.field_op(B.FIELD_FQ2, B.FOP_MUL_KARA, A, Bm)) == [mul(x, y) for x, y in zip(a, b)]
    sub2 = lambda u, v: ((u[0] - v[0]) % p, (u[1] - v[1]) % p)
    add2 = lambda u, v: ((u[0] + v[0]) % p, (u[1] + v[1]) % p)
    assert dec(ctx.field_op(B.FIELD_FQ2, B.FOP_MULSUB2, A, Bm)) == \
        [sub2(mul(x, y), mul(y, add2(x, y))) for x, y in zip(a, b)]
    assert dec(ctx.field_op(B.FIELD_FQ2, B.FOP_ADD, A, Bm)) == [((x[0] + y[0]) % p, (x[1] + y[1]) % p) for x, y in zip(a, b)]
    assert dec(ctx.field_op(B.FIELD_FQ2, B.FOP_SUB, A, Bm)) == [((x[0] - y[0]) % p, (x[1] - y[1]) % p) for x, y in zip(a, b)]
    inv = dec(ctx.field_op(B.FIELD_FQ2, B.FOP_INV, A[:50]))
    for x, y in zip(a[:50], inv):
        assert mul(x, y) == (1, 0)
    inv2 = dec(ctx.field_op(B.FIELD_FQ2, B.FOP_INV_GCD, A[:50]))
    assert inv2 == inv
    # edge values through both inverses: 1, p - 1, 2, (p + 1) / 2
    edge = [1, p - 1, 2, (p + 1) // 2, 3, p - 2]
    em = codec.to_mont(edge, p, L)
    for op in (B.FOP_INV, B.FOP_INV_GCD):
        got = codec.from_mont(ctx.field_op(B.FIELD_FQ, op, em), p)
        assert got == [pow(x, -1, p) for x in edge]
