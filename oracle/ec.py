"""Short-Weierstrass group law over Fq and Fq2 (oracle; test infrastructure only).

Restates the group semantics of ark-ec 0.4 `short_weierstrass::{Affine,Projective}`
(third-party, un-vendored; used by /root/reference/cp-groth16/src/prover.rs:86-147
and committer.rs:85-114).  Points are affine tuples (x, y) or None for the point
at infinity, so every value is canonical and comparisons are exact.  a = 0 on all
four groups.
"""
from .params import get_curve


class Fq:
    """Prime field helper working on plain ints."""

    def __init__(self, p):
        self.p = p
        self.zero = 0
        self.one = 1

    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def mul(self, a, b):
        return a * b % self.p

    def sqr(self, a):
        return a * a % self.p

    def inv(self, a):
        return pow(a, -1, self.p)

    def is_zero(self, a):
        return a % self.p == 0

    def from_int(self, k):
        return k % self.p


class Fq2:
    """Fq[u]/(u^2+1); elements are (c0, c1)."""

    def __init__(self, p):
        self.p = p
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a):
        return self.mul(a, a)

    def inv(self, a):
        p = self.p
        n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
        return (a[0] * n % p, (-a[1] * n) % p)

    def is_zero(self, a):
        return a[0] % self.p == 0 and a[1] % self.p == 0

    def from_int(self, k):
        return (k % self.p, 0)


class Group:
    """y^2 = x^3 + b over field F.  Affine arithmetic (exact, canonical)."""

    def __init__(self, F, b, gen, order, name):
        self.F = F
        self.b = b
        self.gen = gen
        self.order = order
        self.name = name

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)) == F.zero

    def neg(self, P):
        if P is None:
            return None
        return (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if y1 == y2 and not F.is_zero(y1):
                return self.double(P)
            return None
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def double(self, P):
        F = self.F
        if P is None:
            return None
        x1, y1 = P
        if F.is_zero(y1):
            return None
        three_x2 = F.mul(F.from_int(3), F.sqr(x1))
        lam = F.mul(three_x2, F.inv(F.add(y1, y1)))
        x3 = F.sub(F.sqr(lam), F.add(x1, x1))
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def sub(self, P, Q):
        return self.add(P, self.neg(Q))

    # Jacobian helpers (fast scalar multiplication; a = 0)
    def _jdbl(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.mul(F.from_int(8), C)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def _jadd(self, P, Q):
        F = self.F
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        if F.is_zero(Z1):
            return Q
        if F.is_zero(Z2):
            return P
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self._jdbl(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def to_jac(self, P):
        F = self.F
        if P is None:
            return (F.one, F.one, F.zero)
        return (P[0], P[1], F.one)

    def from_jac(self, J):
        """Jacobian (X,Y,Z) -> affine: x = X/Z^2, y = Y/Z^3 (ark-ec into_affine)."""
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def from_xyzz(self, P):
        """XYZZ (X,Y,ZZ,ZZZ) -> affine: x = X/ZZ, y = Y/ZZZ."""
        F = self.F
        X, Y, ZZ, ZZZ = P
        if F.is_zero(ZZ):
            return None
        return (F.mul(X, F.inv(ZZ)), F.mul(Y, F.inv(ZZZ)))

    def mul(self, P, k):
        """k * P for any integer k (taken mod nothing: plain integer multiple)."""
        if P is None or k == 0:
            return None
        if k < 0:
            return self.mul(self.neg(P), -k)
        acc = self.to_jac(None)
        base = self.to_jac(P)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == "1":
                acc = self._jadd(acc, base)
        return self.from_jac(acc)

    def msm_naive(self, bases, scalars):
        """sum_i scalars[i] * bases[i], zip-to-shorter (ark-ec msm_bigint semantics)."""
        acc = self.to_jac(None)
        for P, k in zip(bases, scalars):
            if P is None or k == 0:
                continue
            acc = self._jadd(acc, self.to_jac(self.mul(P, k)))
        return self.from_jac(acc)

    def sum(self, pts):
        acc = self.to_jac(None)
        for P in pts:
            acc = self._jadd(acc, self.to_jac(P))
        return self.from_jac(acc)


_CACHE = {}


def groups(curve):
    """Return (G1, G2) Group objects for a curve name / id / Curve."""
    c = get_curve(curve)
    if c.name not in _CACHE:
        g1 = Group(Fq(c.p), c.b1, c.g1, c.r, c.name + ".G1")
        g2 = Group(Fq2(c.p), c.b2, c.g2, c.r, c.name + ".G2")
        _CACHE[c.name] = (g1, g2)
    return _CACHE[c.name]


# ---------------------------------------------------------------------------
# arkworks-style Pippenger (restated from ark-ec 0.4
# scalar_mul/variable_base/mod.rs: msm_bigint_wnaf + make_digits), used to
# cross-check the C restatement's window rule and digit recoding.
# ---------------------------------------------------------------------------

def ark_ln_without_floats(a):
    # ark_std::log2(a) = ceil(log2(a)); ln ~= log2 * 69 / 100
    log2 = (a - 1).bit_length() if a > 1 else 0
    return log2 * 69 // 100


def ark_window_size(size):
    return 3 if size < 32 else ark_ln_without_floats(size) + 2


def ark_make_digits(a, w, num_bits):
    """Signed base-2^w digits of integer a; the last digit absorbs the carry."""
    radix = 1 << w
    window_mask = radix - 1
    carry = 0
    digits_count = (num_bits + w - 1) // w
    digits = [0] * digits_count
    for i in range(digits_count):
        bit_buf = (a >> (i * w)) & window_mask
        coef = carry + bit_buf
        carry = (coef + radix // 2) >> w
        digits[i] = coef - (carry << w)
    digits[digits_count - 1] += carry << w
    return digits


def msm_pippenger_ark(G, bases, scalars, num_bits):
    size = min(len(bases), len(scalars))
    if size == 0:
        return None
    c = ark_window_size(size)
    digits = [ark_make_digits(s, c, num_bits) for s in scalars[:size]]
    digits_count = (num_bits + c - 1) // c
    window_sums = []
    for i in range(digits_count):
        buckets = [G.to_jac(None)] * (1 << c)
        for d, base in zip(digits, bases):
            s = d[i]
            if base is None:
                continue
            if s > 0:
                buckets[s - 1] = G._jadd(buckets[s - 1], G.to_jac(base))
            elif s < 0:
                buckets[-s - 1] = G._jadd(buckets[-s - 1], G.to_jac(G.neg(base)))
        running = G.to_jac(None)
        res = G.to_jac(None)
        for b in reversed(buckets):
            running = G._jadd(running, b)
            res = G._jadd(res, running)
        window_sums.append(res)
    lowest = window_sums[0]
    total = G.to_jac(None)
    for ws in reversed(window_sums[1:]):
        total = G._jadd(total, ws)
        for _ in range(c):
            total = G._jdbl(total)
    return G.from_jac(G._jadd(lowest, total))
