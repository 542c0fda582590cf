"""Radix-2 evaluation domain + R1CS->QAP witness map (oracle; test infrastructure only).

Restates ark-poly 0.4 `Radix2EvaluationDomain` (domain/radix2/{mod,fft}.rs) and
ark-groth16 0.4 `LibsnarkReduction` (r1cs_to_qap.rs) -- third-party, un-vendored;
the reference reaches them through `QAP::witness_map` at
/root/reference/cp-groth16/src/prover.rs:123 and
`QAP::instance_map_with_evaluation` / `h_query_scalars` at generator.rs:75-76,182.
All values are canonical ints mod r.
"""
from .params import get_curve


class Domain:
    """Radix2EvaluationDomain::new(num_coeffs) (+ get_coset(offset))."""

    def __init__(self, curve, num_coeffs, offset=1):
        c = get_curve(curve)
        self.curve = c
        self.r = c.r
        size = 1
        log = 0
        while size < num_coeffs:
            size <<= 1
            log += 1
        if log > c.two_adicity:
            raise ValueError("PolynomialDegreeTooLarge")
        self.size = size
        self.log_size = log
        self.group_gen = pow(c.two_adic_root, 1 << (c.two_adicity - log), c.r)
        self.group_gen_inv = pow(self.group_gen, -1, c.r)
        self.size_inv = pow(size, -1, c.r)
        self.offset = offset % c.r
        self.offset_inv = pow(self.offset, -1, c.r)
        self.offset_pow_size = pow(self.offset, size, c.r)

    def get_coset(self, offset):
        return Domain(self.curve, self.size, offset)

    # --- transforms (natural order in and out) ---
    def _ntt(self, v, w):
        """out[k] = sum_j v[j] w^{jk}; iterative radix-2, exact."""
        n = self.size
        r = self.r
        a = list(v)
        # bit reversal
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wl = pow(w, n // length, r)
            half = length >> 1
            tw = [1] * half
            for k in range(1, half):
                tw[k] = tw[k - 1] * wl % r
            for s in range(0, n, length):
                for k in range(half):
                    u = a[s + k]
                    t = a[s + k + half] * tw[k] % r
                    a[s + k] = (u + t) % r
                    a[s + k + half] = (u - t) % r
            length <<= 1
        return a

    def fft(self, coeffs):
        """fft_in_place: pad to size, pre-multiply by offset^i on a coset."""
        v = [x % self.r for x in coeffs] + [0] * (self.size - len(coeffs))
        assert len(v) == self.size
        if self.offset != 1:
            g = 1
            for i in range(self.size):
                v[i] = v[i] * g % self.r
                g = g * self.offset % self.r
        return self._ntt(v, self.group_gen)

    def ifft(self, evals):
        """ifft_in_place: inverse transform, * size^-1, and * offset^-i on a coset."""
        v = [x % self.r for x in evals] + [0] * (self.size - len(evals))
        assert len(v) == self.size
        out = self._ntt(v, self.group_gen_inv)
        g = self.size_inv
        for i in range(self.size):
            out[i] = out[i] * g % self.r
            if self.offset != 1:
                g = g * self.offset_inv % self.r
        return out

    def dft_naive(self, coeffs):
        """O(N^2) evaluation at offset * w^k -- the definition the fast path must match."""
        r = self.r
        v = list(coeffs) + [0] * (self.size - len(coeffs))
        out = []
        for k in range(self.size):
            x = self.offset * pow(self.group_gen, k, r) % r
            acc = 0
            for cf in reversed(v):
                acc = (acc * x + cf) % r
            out.append(acc)
        return out

    def evaluate_vanishing_polynomial(self, tau):
        return (pow(tau, self.size, self.r) - self.offset_pow_size) % self.r

    def elements(self):
        x = self.offset
        for _ in range(self.size):
            yield x
            x = x * self.group_gen % self.r

    def evaluate_all_lagrange_coefficients(self, tau):
        """L_i(tau) for the (offset = 1) domain, by the closed form; tau outside the domain."""
        r = self.r
        n = self.size
        assert self.offset == 1
        zt = (pow(tau, n, r) - 1) % r
        assert zt != 0
        out = []
        w = 1
        for _ in range(n):
            # L_i(tau) = Z(tau) * w^i / (n * (tau - w^i))
            out.append(zt * w % r * pow(n * (tau - w) % r, -1, r) % r)
            w = w * self.group_gen % r
        return out


class R1CS:
    """Constraint matrices as the reference sees them after `cs.finalize()` +
    `to_matrices()`: rows are sparse lists of (coeff, column); columns index
    [instance (incl. the constant one at 0) || witness]."""

    def __init__(self, num_instance, num_witness, a_rows, b_rows, c_rows):
        self.num_instance = num_instance
        self.num_witness = num_witness
        self.a = a_rows
        self.b = b_rows
        self.c = c_rows
        assert len(a_rows) == len(b_rows) == len(c_rows)

    @property
    def num_constraints(self):
        return len(self.a)

    @property
    def num_variables(self):
        return self.num_instance + self.num_witness

    def is_satisfied(self, z, r):
        for ra, rb, rc in zip(self.a, self.b, self.c):
            ea = sum(cf * z[i] for cf, i in ra) % r
            eb = sum(cf * z[i] for cf, i in rb) % r
            ec = sum(cf * z[i] for cf, i in rc) % r
            if ea * eb % r != ec:
                return False
        return True


def qap_evals(curve, r1cs, z):
    """The a, b, c evaluation vectors fed to the 7-NTT pipeline
    (ark-groth16 witness_map_from_matrices, first half)."""
    c = get_curve(curve)
    r = c.r
    n = r1cs.num_constraints
    l = r1cs.num_instance
    dom = Domain(c, n + l)
    N = dom.size
    a = [0] * N
    b = [0] * N
    cc = [0] * N
    for i in range(n):
        a[i] = sum(cf * z[j] for cf, j in r1cs.a[i]) % r
        b[i] = sum(cf * z[j] for cf, j in r1cs.b[i]) % r
        cc[i] = sum(cf * z[j] for cf, j in r1cs.c[i]) % r
    for i in range(l):
        a[n + i] = z[i] % r
    return dom, a, b, cc


def h_from_evals(curve, dom, a, b, cc):
    """ifft x3, coset fft x3, (a*b - c) / Z_H(g), coset ifft  (second half of
    witness_map_from_matrices).  Returns all N coefficients."""
    c = get_curve(curve)
    r = c.r
    coset = dom.get_coset(c.fr_generator)
    a = coset.fft(dom.ifft(a))
    b = coset.fft(dom.ifft(b))
    cc = coset.fft(dom.ifft(cc))
    zinv = pow(coset.evaluate_vanishing_polynomial_on_coset(), -1, r)
    ab = [((x * y - w) % r) * zinv % r for x, y, w in zip(a, b, cc)]
    return coset.ifft(ab)


def _vanishing_on_coset(self):
    # Z_H(g * w^k) = g^N - 1 for every k
    return (self.offset_pow_size - 1) % self.r


Domain.evaluate_vanishing_polynomial_on_coset = _vanishing_on_coset


def witness_map(curve, r1cs, z):
    """LibsnarkReduction::witness_map -> h, all N coefficients (h[N-1] == 0)."""
    dom, a, b, cc = qap_evals(curve, r1cs, z)
    return h_from_evals(curve, dom, a, b, cc)


def instance_map_with_evaluation(curve, r1cs, t):
    """LibsnarkReduction::instance_map_with_evaluation: (a, b, c, zt, qap_num_variables, m_raw)."""
    c = get_curve(curve)
    r = c.r
    n = r1cs.num_constraints
    l = r1cs.num_instance
    dom = Domain(c, n + l)
    zt = dom.evaluate_vanishing_polynomial(t)
    u = dom.evaluate_all_lagrange_coefficients(t)
    nv = r1cs.num_variables
    a = [0] * nv
    b = [0] * nv
    cc = [0] * nv
    for i in range(l):
        a[i] = u[n + i]
    for i in range(n):
        for cf, j in r1cs.a[i]:
            a[j] = (a[j] + u[i] * cf) % r
        for cf, j in r1cs.b[i]:
            b[j] = (b[j] + u[i] * cf) % r
        for cf, j in r1cs.c[i]:
            cc[j] = (cc[j] + u[i] * cf) % r
    return a, b, cc, zt, (l - 1) + r1cs.num_witness, dom.size


def h_query_scalars(curve, max_power, t, zt, delta_inv):
    r = get_curve(curve).r
    out = []
    s = zt * delta_inv % r
    x = 1
    for _ in range(max_power):
        out.append(s * x % r)
        x = x * t % r
    return out
