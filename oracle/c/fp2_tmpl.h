/* Fq2 = Fq[u]/(u^2+1) over a base field instantiated by field_tmpl.h:
 *   #define FP2  name prefix of the extension (e.g. bls_fq2)
 *   #define FPB  name prefix of the base field
 * ORACLE / CPU BASELINE ONLY. */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define F(name) CAT(FP2, name)
#define B(name) CAT(FPB, name)

typedef struct { B(t) c0, c1; } F(t);
static F(t) F(ONE);

static inline int F(is_zero)(const F(t) * a) { return B(is_zero)(&a->c0) && B(is_zero)(&a->c1); }
static inline int F(eq)(const F(t) * a, const F(t) * b) { return B(eq)(&a->c0, &b->c0) && B(eq)(&a->c1, &b->c1); }
static inline void F(add)(F(t) * r, const F(t) * a, const F(t) * b) { B(add)(&r->c0, &a->c0, &b->c0); B(add)(&r->c1, &a->c1, &b->c1); }
static inline void F(sub)(F(t) * r, const F(t) * a, const F(t) * b) { B(sub)(&r->c0, &a->c0, &b->c0); B(sub)(&r->c1, &a->c1, &b->c1); }
static inline void F(dbl)(F(t) * r, const F(t) * a) { F(add)(r, a, a); }
static inline void F(neg)(F(t) * r, const F(t) * a) { B(neg)(&r->c0, &a->c0); B(neg)(&r->c1, &a->c1); }
static inline void F(mul)(F(t) * r, const F(t) * a, const F(t) * b) {
    B(t) t0, t1, s0, s1, m;
    B(mul)(&t0, &a->c0, &b->c0);
    B(mul)(&t1, &a->c1, &b->c1);
    B(add)(&s0, &a->c0, &a->c1);
    B(add)(&s1, &b->c0, &b->c1);
    B(mul)(&m, &s0, &s1);
    B(sub)(&r->c0, &t0, &t1);
    B(sub)(&m, &m, &t0);
    B(sub)(&r->c1, &m, &t1);
}
static inline void F(sqr)(F(t) * r, const F(t) * a) {
    B(t) s, d, m;
    B(add)(&s, &a->c0, &a->c1);
    B(sub)(&d, &a->c0, &a->c1);
    B(mul)(&m, &a->c0, &a->c1);
    B(mul)(&r->c0, &s, &d);
    B(dbl)(&r->c1, &m);
}
static void F(inv)(F(t) * r, const F(t) * a) {
    B(t) n, t;
    B(sqr)(&n, &a->c0);
    B(sqr)(&t, &a->c1);
    B(add)(&n, &n, &t);
    B(inv)(&n, &n);
    B(mul)(&r->c0, &a->c0, &n);
    B(mul)(&t, &a->c1, &n);
    B(neg)(&r->c1, &t);
}
static void F(init)(void) { F(ONE).c0 = B(ONE); memset(&F(ONE).c1, 0, sizeof(B(t))); }
#undef F
#undef B
#undef CAT
#undef CAT_
