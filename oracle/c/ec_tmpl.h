/* Short-Weierstrass (a = 0) Jacobian group + arkworks-style Pippenger -- template body:
 *   #define EC   name prefix of the group (e.g. bls_g1)
 *   #define FE   name prefix of the coordinate field (Fq or Fq2 instance)
 * Restates ark-ec 0.4 (third-party, un-vendored): `short_weierstrass::Projective`
 * {double_in_place, add_assign, add_assign_mixed (EFD dbl-2009-l, add-2007-bl, madd-2007-bl)}
 * and `VariableBaseMSM::msm_bigint_wnaf` + `make_digits` (scalar_mul/variable_base/mod.rs):
 * c = 3 if n < 32 else ln_without_floats(n) + 2; signed digits; ONE TASK PER WINDOW (the only
 * parallelism arkworks has inside an MSM); buckets are Projective, bases affine (mixed add);
 * running-sum reduction; windows folded high to low with c doublings.
 * Reached from /root/reference/cp-groth16/src/prover.rs:117,129,168 and committer.rs:89,113.
 * ORACLE / CPU BASELINE ONLY. */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define G(name) CAT(EC, name)
#define E(name) CAT(FE, name)

typedef struct { E(t) x, y; int inf; } G(aff);
typedef struct { E(t) x, y, z; } G(jac);

static inline void G(set_inf)(G(jac) * p) { p->x = E(ONE); p->y = E(ONE); memset(&p->z, 0, sizeof p->z); }
static inline int G(is_inf)(const G(jac) * p) { return E(is_zero)(&p->z); }

static void G(dbl)(G(jac) * p) {
    if (G(is_inf)(p)) return;
    E(t) a, b, c, d, e, f, t;
    E(sqr)(&a, &p->x);
    E(sqr)(&b, &p->y);
    E(sqr)(&c, &b);
    E(add)(&t, &p->x, &b);
    E(sqr)(&t, &t);
    E(sub)(&t, &t, &a);
    E(sub)(&t, &t, &c);
    E(dbl)(&d, &t);
    E(dbl)(&e, &a);
    E(add)(&e, &e, &a);
    E(sqr)(&f, &e);
    E(mul)(&t, &p->y, &p->z);
    E(dbl)(&p->z, &t);
    E(dbl)(&t, &d);
    E(sub)(&p->x, &f, &t);
    E(sub)(&t, &d, &p->x);
    E(mul)(&t, &e, &t);
    E(dbl)(&c, &c);
    E(dbl)(&c, &c);
    E(dbl)(&c, &c);
    E(sub)(&p->y, &t, &c);
}

static void G(add_mixed)(G(jac) * p, const G(aff) * q) {
    if (q->inf) return;
    if (G(is_inf)(p)) { p->x = q->x; p->y = q->y; p->z = E(ONE); return; }
    E(t) z1z1, u2, s2, h, hh, i, j, r, v, t;
    E(sqr)(&z1z1, &p->z);
    E(mul)(&u2, &q->x, &z1z1);
    E(mul)(&s2, &q->y, &p->z);
    E(mul)(&s2, &s2, &z1z1);
    if (E(eq)(&p->x, &u2) && E(eq)(&p->y, &s2)) { G(dbl)(p); return; }
    E(sub)(&h, &u2, &p->x);
    E(sqr)(&hh, &h);
    E(dbl)(&i, &hh);
    E(dbl)(&i, &i);
    E(mul)(&j, &h, &i);
    E(sub)(&r, &s2, &p->y);
    E(dbl)(&r, &r);
    E(mul)(&v, &p->x, &i);
    E(sqr)(&t, &r);
    E(sub)(&t, &t, &j);
    E(sub)(&t, &t, &v);
    E(t) x3;
    E(sub)(&x3, &t, &v);
    E(mul)(&j, &p->y, &j);
    E(dbl)(&j, &j);
    E(sub)(&t, &v, &x3);
    E(mul)(&t, &r, &t);
    E(sub)(&p->y, &t, &j);
    E(add)(&t, &p->z, &h);
    E(sqr)(&t, &t);
    E(sub)(&t, &t, &z1z1);
    E(sub)(&p->z, &t, &hh);
    p->x = x3;
}

static void G(add)(G(jac) * p, const G(jac) * q) {
    if (G(is_inf)(q)) return;
    if (G(is_inf)(p)) { *p = *q; return; }
    E(t) z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
    E(sqr)(&z1z1, &p->z);
    E(sqr)(&z2z2, &q->z);
    E(mul)(&u1, &p->x, &z2z2);
    E(mul)(&u2, &q->x, &z1z1);
    E(mul)(&s1, &p->y, &q->z);
    E(mul)(&s1, &s1, &z2z2);
    E(mul)(&s2, &q->y, &p->z);
    E(mul)(&s2, &s2, &z1z1);
    if (E(eq)(&u1, &u2) && E(eq)(&s1, &s2)) { G(dbl)(p); return; }
    E(sub)(&h, &u2, &u1);
    E(dbl)(&i, &h);
    E(sqr)(&i, &i);
    E(mul)(&j, &h, &i);
    E(sub)(&r, &s2, &s1);
    E(dbl)(&r, &r);
    E(mul)(&v, &u1, &i);
    E(t) x3;
    E(sqr)(&t, &r);
    E(sub)(&t, &t, &j);
    E(sub)(&t, &t, &v);
    E(sub)(&x3, &t, &v);
    E(mul)(&j, &s1, &j);
    E(dbl)(&j, &j);
    E(sub)(&t, &v, &x3);
    E(mul)(&t, &r, &t);
    E(sub)(&p->y, &t, &j);
    E(add)(&t, &p->z, &q->z);
    E(sqr)(&t, &t);
    E(sub)(&t, &t, &z1z1);
    E(sub)(&t, &t, &z2z2);
    E(mul)(&p->z, &t, &h);
    p->x = x3;
}

static void G(to_affine)(G(aff) * r, const G(jac) * p) {
    if (G(is_inf)(p)) { memset(r, 0, sizeof *r); r->inf = 1; return; }
    E(t) zi, zi2;
    E(inv)(&zi, &p->z);
    E(sqr)(&zi2, &zi);
    E(mul)(&r->x, &p->x, &zi2);
    E(mul)(&zi2, &zi2, &zi);
    E(mul)(&r->y, &p->y, &zi2);
    r->inf = 0;
}

/* k * P, k canonical little-endian u64[4] (double-and-add, MSB first) */
static void G(mul_u256)(G(jac) * r, const G(aff) * p, const uint64_t k[4]) {
    G(jac) acc;
    G(set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        G(dbl)(&acc);
        if ((k[i >> 6] >> (i & 63)) & 1) G(add_mixed)(&acc, p);
    }
    *r = acc;
}

/* msm_bigint_wnaf.  scalars: n x 4 u64 canonical.  threads: OpenMP threads over windows. */
static void G(msm)(G(jac) * out, const G(aff) * bases, const uint64_t* scalars, size_t n, int num_bits,
                   int threads) {
    G(set_inf)(out);
    if (n == 0) return;
    int c = ark_window_size(n);
    int digits_count = (num_bits + c - 1) / c;
    int32_t* digits = (int32_t*)malloc(sizeof(int32_t) * n * (size_t)digits_count);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) ark_make_digits(scalars + 4 * i, c, num_bits, digits + i * (size_t)digits_count);
    G(jac)* window_sums = (G(jac)*)malloc(sizeof(G(jac)) * digits_count);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int w = 0; w < digits_count; w++) {
        /* ark-ec allocates 1 << (c-1) buckets; the unrecoded top digit can exceed that for some
         * (c, modulus) pairs, so size by the largest digit actually present (value-neutral). */
        size_t nb = (size_t)1 << (c - 1);
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * (size_t)digits_count + w];
            size_t ad = (size_t)(d < 0 ? -d : d);
            if (ad > nb) nb = ad;
        }
        G(jac)* buckets = (G(jac)*)malloc(sizeof(G(jac)) * nb);
        for (size_t b = 0; b < nb; b++) G(set_inf)(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * (size_t)digits_count + w];
            if (d > 0) {
                G(add_mixed)(&buckets[d - 1], &bases[i]);
            } else if (d < 0) {
                G(aff) neg = bases[i];
                E(neg)(&neg.y, &neg.y);
                G(add_mixed)(&buckets[-d - 1], &neg);
            }
        }
        G(jac) running, res;
        G(set_inf)(&running);
        G(set_inf)(&res);
        for (size_t b = nb; b-- > 0;) {
            G(add)(&running, &buckets[b]);
            G(add)(&res, &running);
        }
        window_sums[w] = res;
        free(buckets);
    }
    G(jac) total;
    G(set_inf)(&total);
    for (int w = digits_count - 1; w >= 1; w--) {
        G(add)(&total, &window_sums[w]);
        for (int k = 0; k < c; k++) G(dbl)(&total);
    }
    G(add)(&total, &window_sums[0]);
    *out = total;
    free(window_sums);
    free(digits);
}
#undef G
#undef E
#undef CAT
#undef CAT_
