/* Montgomery prime field with NL 64-bit limbs -- template body, include with:
 *   #define FP        name prefix (e.g. bls_fq)
 *   #define NL        limb count (4 or 6)
 *   #define MOD_INIT  {limbs...}  modulus, little-endian (compile-time constant)
 *   #define INV_CONST -p^-1 mod 2^64
 * Restates ark-ff 0.4 `Fp<MontBackend<_,N>,N>` (third-party, un-vendored; B.4 of SURVEY.md):
 * little-endian u64 limbs, Montgomery form R = 2^(64 NL), CIOS multiplication.
 * ORACLE / CPU BASELINE ONLY -- test infrastructure, never linked into the product. */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define F(name) CAT(FP, name)

typedef struct { uint64_t v[NL]; } F(t);

static const F(t) F(MOD) = {MOD_INIT};
static F(t) F(ONE), F(R2);
static const uint64_t F(INV) = INV_CONST;     /* -p^-1 mod 2^64 */

static inline int F(is_zero)(const F(t) * a) {
    uint64_t o = 0;
    for (int i = 0; i < NL; i++) o |= a->v[i];
    return o == 0;
}
static inline int F(eq)(const F(t) * a, const F(t) * b) {
    uint64_t o = 0;
    for (int i = 0; i < NL; i++) o |= a->v[i] ^ b->v[i];
    return o == 0;
}
static inline int F(geq_mod)(const F(t) * a) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a->v[i] > F(MOD).v[i]) return 1;
        if (a->v[i] < F(MOD).v[i]) return 0;
    }
    return 1;
}
static inline void F(sub_mod_raw)(F(t) * a) {
    unsigned __int128 bw = 0;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 d = (unsigned __int128)a->v[i] - F(MOD).v[i] - (uint64_t)bw;
        a->v[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
}
static inline void F(add)(F(t) * r, const F(t) * a, const F(t) * b) {
    unsigned __int128 c = 0;
    for (int i = 0; i < NL; i++) {
        c += (unsigned __int128)a->v[i] + b->v[i];
        r->v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (F(geq_mod)(r)) F(sub_mod_raw)(r);
}
static inline void F(sub)(F(t) * r, const F(t) * a, const F(t) * b) {
    unsigned __int128 bw = 0;
    F(t) t;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 d = (unsigned __int128)a->v[i] - b->v[i] - (uint64_t)bw;
        t.v[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        unsigned __int128 c = 0;
        for (int i = 0; i < NL; i++) {
            c += (unsigned __int128)t.v[i] + F(MOD).v[i];
            t.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    *r = t;
}
static inline void F(dbl)(F(t) * r, const F(t) * a) { F(add)(r, a, a); }
static inline void F(neg)(F(t) * r, const F(t) * a) {
    if (F(is_zero)(a)) { *r = *a; return; }
    F(t) z; memset(&z, 0, sizeof z);
    F(sub)(r, &z, a);
}
/* Montgomery product.  On x86-64 with BMI2 + ADX: the MULX / ADCX / ADOX routine of mont_asm.h (the instruction
 * mix of ark-ff's `asm` feature, which the reference enables by default); otherwise portable CIOS (fully unrolled;
 * branch-free final subtraction). */
static inline __attribute__((always_inline)) void F(mul)(F(t) * r, const F(t) * a, const F(t) * b) {
#ifdef ORC_HAVE_MONT_ASM
    if (g_orc_use_asm) {
        if (NL == 6) orc_mont_mul_asm_6(r->v, a->v, b->v, F(MOD).v, F(INV));
        else orc_mont_mul_asm_4(r->v, a->v, b->v, F(MOD).v, F(INV));
        return;
    }
#endif
    uint64_t t[NL + 2] = {0};
#pragma GCC unroll 8
    for (int i = 0; i < NL; i++) {
        unsigned __int128 c = 0;
#pragma GCC unroll 8
        for (int j = 0; j < NL; j++) {
            c += (unsigned __int128)a->v[j] * b->v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[NL];
        t[NL] = (uint64_t)c;
        t[NL + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F(INV);
        c = (unsigned __int128)m * F(MOD).v[0] + t[0];
        c >>= 64;
#pragma GCC unroll 8
        for (int j = 1; j < NL; j++) {
            c += (unsigned __int128)m * F(MOD).v[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[NL];
        t[NL - 1] = (uint64_t)c;
        t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
    }
    uint64_t s[NL];
    unsigned __int128 bw = 0;
#pragma GCC unroll 8
    for (int i = 0; i < NL; i++) {
        unsigned __int128 d = (unsigned __int128)t[i] - F(MOD).v[i] - (uint64_t)bw;
        s[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    int ge = (t[NL] != 0) | (bw == 0);
#pragma GCC unroll 8
    for (int i = 0; i < NL; i++) r->v[i] = ge ? s[i] : t[i];
}
static inline void F(sqr)(F(t) * r, const F(t) * a) { F(mul)(r, a, a); }
static inline void F(from_mont)(F(t) * r, const F(t) * a) {
    F(t) one; memset(&one, 0, sizeof one); one.v[0] = 1;
    F(mul)(r, a, &one);
}
static inline void F(to_mont)(F(t) * r, const F(t) * a) { F(mul)(r, a, &F(R2)); }
/* a^(p-2) */
static void F(inv)(F(t) * r, const F(t) * a) {
    F(t) e = F(MOD);
    e.v[0] -= 2;
    F(t) acc = F(ONE);
    int started = 0;
    for (int i = NL * 64 - 1; i >= 0; i--) {
        if (started) F(sqr)(&acc, &acc);
        if ((e.v[i >> 6] >> (i & 63)) & 1) {
            if (started) F(mul)(&acc, &acc, a); else { acc = *a; started = 1; }
        }
    }
    *r = acc;
}
static void F(init)(const uint64_t* mod) {
    for (int i = 0; i < NL; i++) if (F(MOD).v[i] != mod[i]) abort();
    /* check inv = -p^-1 mod 2^64 by Newton iteration */
    uint64_t p0 = mod[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
    if (F(INV) != (uint64_t)(0 - x)) abort();
    /* R mod p by doubling 1, 64*NL times; R2 by another 64*NL doublings */
    F(t) one; memset(&one, 0, sizeof one); one.v[0] = 1;
    F(t) x1 = one;
    for (int i = 0; i < 64 * NL; i++) F(add)(&x1, &x1, &x1);
    F(ONE) = x1;
    for (int i = 0; i < 64 * NL; i++) F(add)(&x1, &x1, &x1);
    F(R2) = x1;
}
#undef F
#undef CAT
#undef CAT_
