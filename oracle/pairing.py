"""Pairing-based proof verification (oracle; TEST INFRASTRUCTURE ONLY).

Restates /root/reference/cp-groth16/src/verifier.rs:
  prepare_verifying_key             verifier.rs:7-18   (alpha_beta_gt, -gamma_h, -deltas_h)
  prepare_inputs                    verifier.rs:47-61  (g_ic = gamma_abc_g[0] + sum input_i * gamma_abc_g[i+1])
  verify_proof_with_prepared_inputs verifier.rs:23-43  (multi Miller loop over
        (A, B), (g_ic, -gamma_h), (D_i, -delta_i_h)..., (C, -delta_last_h); final exponentiation;
        compare with e(alpha_g, beta_h))
  verify_proof                      verifier.rs:64-71

The pairing itself lives in the un-vendored ark-ec 0.4 / ark-bls12-381 / ark-bn254 crates.  verify_proof only
compares two values of the SAME pairing, so its accept / reject decision is the same for every non-degenerate
bilinear pairing on (G1, G2); what is restated here is the reduced ate pairing
        e(P, Q) = f_{T,Q}(P)^((p^12 - 1) / r),   T = t - 1  (BLS12-381: T = x, negative; BN254: T = 6 x^2)
with the published curve parameters x, computed on the sextic twist (slopes in Fq2, line values embedded sparsely
in Fq12 = Fq2[w] / (w^6 - xi)).  It is deliberately the plain textbook algorithm (affine Miller loop, generic Fq12
multiplication, square-and-multiply hard part): an independent check of GPU proofs that shares no code with the
trapdoor known answer.  Pinned by tests/test_oracle_pairing.py: bilinearity in both arguments, non-degeneracy,
e(P, Q)^r = 1, and acceptance / rejection of oracle proofs.
"""
from .ec import groups
from .params import get_curve

# curve parameter x of the BLS12 / BN families (public constants of the two curves)
BLS12_381_X = -0xd201000000010000
BN254_X = 4965661367192848881


class _Tower:
    """Fq12 = Fq2[w] / (w^6 - xi), Fq2 = Fq[u] / (u^2 + 1).  Elements: tuples of six Fq2 pairs."""

    def __init__(self, curve):
        c = get_curve(curve)
        self.c = c
        self.p = c.p
        if c.name == "bls12_381":
            self.xi = (1, 1)                 # M-type twist: b' = b * xi
            self.mtype = True
            self.T = BLS12_381_X
            assert c.b2 == self.m2((c.b1, 0), self.xi)
        else:
            self.xi = (9, 1)                 # D-type twist: b' = b / xi
            self.mtype = False
            self.T = 6 * BN254_X * BN254_X
            assert self.m2(c.b2, self.xi) == (c.b1, 0)
            # BN parametrisation of p and r
            x = BN254_X
            assert c.p == 36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1
            assert c.r == 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1
        self.zero2 = (0, 0)
        self.one = ((1, 0),) + ((0, 0),) * 5
        # w^(p^2) = w * xi^((p^2 - 1) / 6): Frobenius^2 acts on coefficient i by gamma^i (Fq2 is fixed by p^2)
        self.gamma = self.pow2(self.xi, (self.p * self.p - 1) // 6)
        self.hard = (self.p**4 - self.p**2 + 1) // c.r
        assert (self.p**4 - self.p**2 + 1) % c.r == 0

    # ---- Fq2 ----
    def a2(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def s2(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def m2(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def n2(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def i2(self, a):
        p = self.p
        n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
        return (a[0] * n % p, (-a[1] * n) % p)

    def pow2(self, a, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = self.m2(r, a)
            a = self.m2(a, a)
            e >>= 1
        return r

    # ---- Fq12 (flat in w) ----
    def mul(self, f, g):
        p = self.p
        acc = [[0, 0] for _ in range(11)]
        for i, a in enumerate(f):
            if a[0] == 0 and a[1] == 0:
                continue
            for j, b in enumerate(g):
                t = acc[i + j]
                t[0] += a[0] * b[0] - a[1] * b[1]
                t[1] += a[0] * b[1] + a[1] * b[0]
        xi0, xi1 = self.xi
        out = []
        for k in range(6):
            lo = acc[k]
            if k < 5:
                h0, h1 = acc[k + 6]
                lo = (lo[0] + h0 * xi0 - h1 * xi1, lo[1] + h0 * xi1 + h1 * xi0)
            out.append((lo[0] % p, lo[1] % p))
        return tuple(out)

    def sqr(self, f):
        return self.mul(f, f)

    def conj(self, f):
        """f^(p^6): w -> -w."""
        return tuple(c if i % 2 == 0 else self.n2(c) for i, c in enumerate(f))

    def frob2(self, f):
        """f^(p^2)."""
        out, g = [], (1, 0)
        for c in f:
            out.append(self.m2(c, g))
            g = self.m2(g, self.gamma)
        return tuple(out)

    # Fq6 = Fq2[v] / (v^3 - xi), v = w^2: triples of Fq2
    def _m6(self, a, b):
        m2, a2 = self.m2, self.a2
        t0, t1, t2 = m2(a[0], b[0]), m2(a[1], b[1]), m2(a[2], b[2])
        c0 = a2(t0, m2(self.xi, a2(m2(a[1], b[2]), m2(a[2], b[1]))))
        c1 = a2(a2(m2(a[0], b[1]), m2(a[1], b[0])), m2(self.xi, t2))
        c2 = a2(a2(m2(a[0], b[2]), m2(a[2], b[0])), t1)
        return (c0, c1, c2)

    def _i6(self, a):
        m2, s2, a2, xi = self.m2, self.s2, self.a2, self.xi
        c0 = s2(m2(a[0], a[0]), m2(xi, m2(a[1], a[2])))
        c1 = s2(m2(xi, m2(a[2], a[2])), m2(a[0], a[1]))
        c2 = s2(m2(a[1], a[1]), m2(a[0], a[2]))
        t = a2(m2(a[0], c0), m2(xi, a2(m2(a[2], c1), m2(a[1], c2))))
        ti = self.i2(t)
        return (m2(c0, ti), m2(c1, ti), m2(c2, ti))

    def inv(self, f):
        """(a + b w)^-1 = (a - b w) / (a^2 - b^2 v)   with a = even, b = odd coefficients (in v = w^2)."""
        a = (f[0], f[2], f[4])
        b = (f[1], f[3], f[5])
        aa = self._m6(a, a)
        bb = self._m6(b, b)
        bbv = (self.m2(self.xi, bb[2]), bb[0], bb[1])            # * v
        d = self._i6(tuple(self.s2(x, y) for x, y in zip(aa, bbv)))
        ra = self._m6(a, d)
        rb = self._m6(b, d)
        return (ra[0], self.n2(rb[0]), ra[1], self.n2(rb[1]), ra[2], self.n2(rb[2]))

    def pow(self, f, e):
        r = self.one
        for bit in bin(e)[2:]:
            r = self.sqr(r)
            if bit == "1":
                r = self.mul(r, f)
        return r

    # ---- Miller loop on the twist ----
    def _line(self, lam, T, P):
        """Value at P = (xP, yP) in E(Fq) of the line of slope `lam` (twist coordinates) through the twist point
        T, up to a factor in a proper subfield (removed by the final exponentiation).
        D-type (untwist x = x' w^2, y = y' w^3, slope lam w):  yP - lam xP w + (lam xT - yT) w^3
        M-type (untwist x = x'/w^2, y = y'/w^3, slope lam/w), times w^3:  (lam xT - yT) - lam xP w^2 + yP w^3"""
        xP, yP = P
        c = self.s2(self.m2(lam, T[0]), T[1])
        lx = ((-lam[0] * xP) % self.p, (-lam[1] * xP) % self.p)
        z = self.zero2
        if self.mtype:
            return (c, z, lx, (yP % self.p, 0), z, z)
        return ((yP % self.p, 0), lx, z, c, z, z)

    def _vertical_free_step(self, f, T, Q2, P, G2):
        """One addition step T <- T + Q2 with the line through them (vertical lines lie in a proper subfield)."""
        if T[0] == Q2[0]:
            if T[1] == Q2[1]:
                lam = self.m2(self.m2((3, 0), self.m2(T[0], T[0])), self.i2(self.a2(T[1], T[1])))
            else:
                return f, None                               # T + (-T): vertical line only
        else:
            lam = self.m2(self.s2(Q2[1], T[1]), self.i2(self.s2(Q2[0], T[0])))
        f = self.mul(f, self._line(lam, T, P))
        return f, G2.add(T, Q2)

    def miller_loop(self, P, Q):
        """f_{|T|,Q}(P) (conjugated when T < 0).  P in G1 (affine ints), Q in G2 (affine Fq2 pairs)."""
        if P is None or Q is None:
            return self.one
        G2 = groups(self.c)[1]
        n = abs(self.T)
        f = self.one
        T = Q
        for bit in bin(n)[3:]:
            lam = self.m2(self.m2((3, 0), self.m2(T[0], T[0])), self.i2(self.a2(T[1], T[1])))
            f = self.mul(self.sqr(f), self._line(lam, T, P))
            T = G2.double(T)
            if bit == "1":
                f, T = self._vertical_free_step(f, T, Q, P, G2)
        if self.T < 0:
            f = self.conj(f)      # f_{-n} = 1 / f_n up to subfield factors; inverse = conjugate after the easy part
        return f

    def final_exponentiation(self, f):
        """f^((p^12 - 1) / r) = ((f^(p^6 - 1))^(p^2 + 1))^((p^4 - p^2 + 1) / r); None for f = 0."""
        if all(c == (0, 0) for c in f):
            return None
        t = self.mul(self.conj(f), self.inv(f))
        t = self.mul(self.frob2(t), t)
        return self.pow(t, self.hard)


_TOWERS = {}


def tower(curve):
    c = get_curve(curve)
    if c.name not in _TOWERS:
        _TOWERS[c.name] = _Tower(c)
    return _TOWERS[c.name]


def multi_miller_loop(curve, g1_points, g2_points):
    """E::multi_miller_loop: product of the Miller loops of the pairs (zip semantics)."""
    t = tower(curve)
    f = t.one
    for P, Q in zip(g1_points, g2_points):
        f = t.mul(f, t.miller_loop(P, Q))
    return f


def final_exponentiation(curve, f):
    return tower(curve).final_exponentiation(f)


def pairing(curve, P, Q):
    """E::pairing(P, Q) -> an element of GT (tuple of six Fq2 pairs)."""
    return final_exponentiation(curve, multi_miller_loop(curve, [P], [Q]))


def gt_pow(curve, g, e):
    return tower(curve).pow(g, e)


def gt_one(curve):
    return tower(curve).one


# ---- verifier.rs ------------------------------------------------------------------------------------
class PreparedVerifyingKey:
    """data_structures.rs:49-60."""

    def __init__(self, vk, alpha_beta_gt, neg_gamma_h, neg_deltas_h):
        self.vk = vk
        self.alpha_beta_gt = alpha_beta_gt
        self.neg_gamma_h = neg_gamma_h
        self.neg_deltas_h = neg_deltas_h


def prepare_verifying_key(vk):
    """verifier.rs:7-18.  `vk`: anything with .curve, .alpha_g, .beta_h, .gamma_h, .gamma_abc_g, .deltas_h
    (oracle.cpgroth16.ProvingKey carries these fields)."""
    G2 = groups(vk.curve)[1]
    return PreparedVerifyingKey(vk, pairing(vk.curve, vk.alpha_g, vk.beta_h), G2.neg(vk.gamma_h),
                                [G2.neg(d) for d in vk.deltas_h])


def prepare_inputs(pvk, public_inputs):
    """verifier.rs:47-61; public_inputs exclude the leading one."""
    vk = pvk.vk
    if len(public_inputs) + 1 != len(vk.gamma_abc_g):
        raise ValueError("MalformedVerifyingKey")
    G1 = groups(vk.curve)[0]
    g_ic = vk.gamma_abc_g[0]
    for x, base in zip(public_inputs, vk.gamma_abc_g[1:]):
        g_ic = G1.add(g_ic, G1.mul(base, x % get_curve(vk.curve).r))
    return g_ic


def verify_proof_with_prepared_inputs(pvk, a, b, c, ds, prepared_inputs):
    """verifier.rs:23-43.  a, c, ds in G1, b in G2 (affine; None = infinity)."""
    curve = pvk.vk.curve
    lhs = [a, prepared_inputs] + list(ds) + [c]
    rhs = [b, pvk.neg_gamma_h] + list(pvk.neg_deltas_h)
    test = final_exponentiation(curve, multi_miller_loop(curve, lhs, rhs))
    if test is None:
        raise ValueError("UnexpectedIdentity")
    return test == pvk.alpha_beta_gt


def verify_proof(pvk, a, b, c, ds, public_inputs):
    """verifier.rs:64-71."""
    return verify_proof_with_prepared_inputs(pvk, a, b, c, ds, prepare_inputs(pvk, public_inputs))
