/* CPU oracle / CPU baseline for the Hekaton CP-Groth16 hot path -- plain C, 64-bit limbs.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  It restates the algorithms the reference
 * executes on its CPU path (arkworks 0.4 semantics; the crates are third-party and not vendored
 * in /root/reference):
 *   orc_msm_*          ark-ec VariableBaseMSM::msm_bigint (wNAF Pippenger, window-parallel)
 *   orc_ntt            ark-poly Radix2EvaluationDomain::{fft,ifft}_in_place (+ coset)
 *   orc_witness_map    ark-groth16 LibsnarkReduction::witness_map_from_matrices
 *   orc_commit/prove   /root/reference/cp-groth16/src/committer.rs:55-122, prover.rs:53-171
 * PARITY: unpinned against the reference's own tests (it has no numeric fixtures); pinned against
 * the Python big-integer oracle (tests/test_oracle_c.py) which is itself pinned against public
 * curve constants and the trapdoor known-answer.
 *
 * All field elements cross this API in Montgomery form, little-endian u64 limbs (the ark-ff
 * memory layout), exactly like include/hekaton_b200.h, so the same buffers feed both sides. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mont_asm.h"
/* 1 when the MULX/ADCX/ADOX multiplication may be used (decided once in init_all; ORC_NO_ASM=1 disables it) */
static int g_orc_use_asm = 0;

/* ---- arkworks window rule and signed-digit recoding ------------------------------------------ */
static int ark_log2_ceil(size_t a) { int l = 0; while (((size_t)1 << l) < a) l++; return l; }
static int ark_window_size(size_t n) { return n < 32 ? 3 : ark_log2_ceil(n) * 69 / 100 + 2; }
static void ark_make_digits(const uint64_t a[4], int w, int num_bits, int32_t* digits) {
    uint64_t radix = 1ull << w, mask = radix - 1, carry = 0;
    int count = (num_bits + w - 1) / w;
    for (int i = 0; i < count; i++) {
        int bit = i * w, limb = bit >> 6, sh = bit & 63;
        uint64_t buf = 0;
        if (limb < 4) {
            buf = a[limb] >> sh;
            if (sh + w > 64 && limb + 1 < 4) buf |= a[limb + 1] << (64 - sh);
        }
        uint64_t coef = carry + (buf & mask);
        carry = (coef + radix / 2) >> w;
        int64_t d = (int64_t)coef - (int64_t)(carry << w);
        digits[i] = (int32_t)d;
    }
    digits[count - 1] += (int32_t)(carry << w);
}

/* ---- field / group instantiations -------------------------------------------------------------- */
#define FP bls_fq
#define NL 6
#define MOD_INIT {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull}
#define INV_CONST 0x89f3fffcfffcfffdull
#include "field_tmpl.h"
#undef FP
#undef NL
#undef MOD_INIT
#undef INV_CONST
#define FP bls_fr
#define NL 4
#define MOD_INIT {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull}
#define INV_CONST 0xfffffffeffffffffull
#include "field_tmpl.h"
#undef FP
#undef NL
#undef MOD_INIT
#undef INV_CONST
#define FP bn_fq
#define NL 4
#define MOD_INIT {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}
#define INV_CONST 0x87d20782e4866389ull
#include "field_tmpl.h"
#undef FP
#undef NL
#undef MOD_INIT
#undef INV_CONST
#define FP bn_fr
#define NL 4
#define MOD_INIT {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}
#define INV_CONST 0xc2e1f593efffffffull
#include "field_tmpl.h"
#undef FP
#undef NL
#undef MOD_INIT
#undef INV_CONST
#define FP2 bls_fq2
#define FPB bls_fq
#include "fp2_tmpl.h"
#undef FP2
#undef FPB
#define FP2 bn_fq2
#define FPB bn_fq
#include "fp2_tmpl.h"
#undef FP2
#undef FPB
#define EC bls_g1
#define FE bls_fq
#include "ec_tmpl.h"
#undef EC
#undef FE
#define EC bls_g2
#define FE bls_fq2
#include "ec_tmpl.h"
#undef EC
#undef FE
#define EC bn_g1
#define FE bn_fq
#include "ec_tmpl.h"
#undef EC
#undef FE
#define EC bn_g2
#define FE bn_fq2
#include "ec_tmpl.h"
#undef EC
#undef FE

static const uint64_t BLS_P[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                                  0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t BLS_R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                  0x73eda753299d7d48ull};
static const uint64_t BN_P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull,
                                 0x30644e72e131a029ull};
static const uint64_t BN_R[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                                 0x30644e72e131a029ull};

static int g_init = 0;
static void init_all(void) {
    if (g_init) return;
#if defined(ORC_HAVE_MONT_ASM)
    __builtin_cpu_init();
    g_orc_use_asm = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx") && !getenv("ORC_NO_ASM");
#endif
    bls_fq_init(BLS_P);
    bls_fr_init(BLS_R);
    bn_fq_init(BN_P);
    bn_fr_init(BN_R);
    bls_fq2_init();
    bn_fq2_init();
    g_init = 1;
}

/* ---- Fr-generic NTT / witness map (both scalar fields have 4 limbs) ---------------------------- */
typedef struct {
    void (*mul)(void*, const void*, const void*);
    void (*add)(void*, const void*, const void*);
    void (*sub)(void*, const void*, const void*);
    void (*inv)(void*, const void*);
    uint64_t one[4];
    uint64_t root[4];      /* TWO_ADIC_ROOT_OF_UNITY, Montgomery */
    uint64_t gen[4];       /* F::GENERATOR, Montgomery */
    int two_adicity;
    int bits;
} fr_ops;

typedef struct { uint64_t v[4]; } fr_t;

#define FR_WRAP(P)                                                                                   \
    static void P##_mul_v(void* r, const void* a, const void* b) { P##_mul((P##_t*)r, (const P##_t*)a, (const P##_t*)b); } \
    static void P##_add_v(void* r, const void* a, const void* b) { P##_add((P##_t*)r, (const P##_t*)a, (const P##_t*)b); } \
    static void P##_sub_v(void* r, const void* a, const void* b) { P##_sub((P##_t*)r, (const P##_t*)a, (const P##_t*)b); } \
    static void P##_inv_v(void* r, const void* a) { P##_inv((P##_t*)r, (const P##_t*)a); }
FR_WRAP(bls_fr)
FR_WRAP(bn_fr)

static void fr_pow_u64(const fr_ops* f, fr_t* r, const fr_t* a, uint64_t e) {
    fr_t acc, base = *a;
    memcpy(&acc, f->one, 32);
    while (e) {
        if (e & 1) f->mul(&acc, &acc, &base);
        f->mul(&base, &base, &base);
        e >>= 1;
    }
    *r = acc;
}

static void get_fr(int curve, fr_ops* f) {
    init_all();
    if (curve == 0) {
        f->mul = bls_fr_mul_v; f->add = bls_fr_add_v; f->sub = bls_fr_sub_v; f->inv = bls_fr_inv_v;
        memcpy(f->one, bls_fr_ONE.v, 32);
        f->two_adicity = 32; f->bits = 255;
        /* GENERATOR = 7; root = 7^((r-1)/2^32) */
        fr_t g; memset(&g, 0, 32);
        fr_t one; memcpy(&one, f->one, 32);
        for (int i = 0; i < 7; i++) f->add(&g, &g, &one);
        memcpy(f->gen, &g, 32);
        /* exponent (r-1) >> 32 as big integer: do square-and-multiply over the 223 high bits */
        uint64_t e[4] = {BLS_R[0] - 1, BLS_R[1], BLS_R[2], BLS_R[3]};
        fr_t acc = one;
        for (int i = 255; i >= 32; i--) {
            f->mul(&acc, &acc, &acc);
            if ((e[i >> 6] >> (i & 63)) & 1) f->mul(&acc, &acc, &g);
        }
        memcpy(f->root, &acc, 32);
    } else {
        f->mul = bn_fr_mul_v; f->add = bn_fr_add_v; f->sub = bn_fr_sub_v; f->inv = bn_fr_inv_v;
        memcpy(f->one, bn_fr_ONE.v, 32);
        f->two_adicity = 28; f->bits = 254;
        fr_t g; memset(&g, 0, 32);
        fr_t one; memcpy(&one, f->one, 32);
        for (int i = 0; i < 5; i++) f->add(&g, &g, &one);
        memcpy(f->gen, &g, 32);
        uint64_t e[4] = {BN_R[0] - 1, BN_R[1], BN_R[2], BN_R[3]};
        fr_t acc = one;
        for (int i = 255; i >= 28; i--) {
            f->mul(&acc, &acc, &acc);
            if ((e[i >> 6] >> (i & 63)) & 1) f->mul(&acc, &acc, &g);
        }
        memcpy(f->root, &acc, 32);
    }
}

static uint32_t bitrev32(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

/* In-order radix-2 transform out[k] = sum_j x_j w^(jk) (bit-reverse, then DIT rounds with a
 * per-round root table -- the shape of ark-poly's serial/parallel radix-2 kernel). */
static void fr_ntt_core(const fr_ops* f, fr_t* x, int log_n, const fr_t* w, int threads) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev32((uint32_t)i, log_n);
        if (i < j) { fr_t t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    fr_t* tw = (fr_t*)malloc(sizeof(fr_t) * (n / 2 + 1));
    for (int s = 1; s <= log_n; s++) {
        size_t len = (size_t)1 << s, half = len >> 1;
        fr_t wl;
        fr_pow_u64(f, &wl, w, (uint64_t)(n / len));
        memcpy(&tw[0], f->one, 32);
        for (size_t k = 1; k < half; k++) f->mul(&tw[k], &tw[k - 1], &wl);
        size_t nblocks = n / len;
        if (nblocks >= (size_t)threads * 4) {
#pragma omp parallel for num_threads(threads) schedule(static)
            for (size_t b = 0; b < nblocks; b++) {
                fr_t* p = x + b * len;
                for (size_t k = 0; k < half; k++) {
                    fr_t t, u = p[k];
                    f->mul(&t, &p[k + half], &tw[k]);
                    f->add(&p[k], &u, &t);
                    f->sub(&p[k + half], &u, &t);
                }
            }
        } else {
            for (size_t b = 0; b < nblocks; b++) {
                fr_t* p = x + b * len;
#pragma omp parallel for num_threads(threads) schedule(static)
                for (size_t k = 0; k < half; k++) {
                    fr_t t, u = p[k];
                    f->mul(&t, &p[k + half], &tw[k]);
                    f->add(&p[k], &u, &t);
                    f->sub(&p[k + half], &u, &t);
                }
            }
        }
    }
    free(tw);
}

static void fr_domain_root(const fr_ops* f, int log_n, fr_t* w) {
    fr_t r; memcpy(&r, f->root, 32);
    for (int i = log_n; i < f->two_adicity; i++) f->mul(&r, &r, &r);
    *w = r;
}

/* dir 0 = fft_in_place, 1 = ifft_in_place; coset 0/1 (offset = GENERATOR) */
static void fr_ntt(const fr_ops* f, fr_t* x, int log_n, int dir, int coset, int threads) {
    size_t n = (size_t)1 << log_n;
    fr_t w, g; fr_domain_root(f, log_n, &w);
    memcpy(&g, f->gen, 32);
    if (dir == 0) {
        if (coset) {
            fr_t p; memcpy(&p, f->one, 32);
            for (size_t i = 0; i < n; i++) { f->mul(&x[i], &x[i], &p); f->mul(&p, &p, &g); }
        }
        fr_ntt_core(f, x, log_n, &w, threads);
    } else {
        fr_t wi; f->inv(&wi, &w);
        fr_ntt_core(f, x, log_n, &wi, threads);
        fr_t nn; memcpy(&nn, f->one, 32);
        for (int i = 0; i < log_n; i++) f->add(&nn, &nn, &nn);
        fr_t ninv; f->inv(&ninv, &nn);
        if (!coset) {
#pragma omp parallel for num_threads(threads) schedule(static)
            for (size_t i = 0; i < n; i++) f->mul(&x[i], &x[i], &ninv);
        } else {
            fr_t gi; f->inv(&gi, &g);
            fr_t p = ninv;
            for (size_t i = 0; i < n; i++) { f->mul(&x[i], &x[i], &p); f->mul(&p, &p, &gi); }
        }
    }
}

int orc_ntt(int curve, uint64_t* data, int log_n, int dir, int coset, int threads) {
    fr_ops f; get_fr(curve, &f);
    if (log_n > f.two_adicity) return -6;
    if (threads < 1) threads = 1;
    fr_ntt(&f, (fr_t*)data, log_n, dir, coset, threads);
    return 0;
}

/* witness_map_from_matrices.  CSR matrices (ptr n+1, col, val Montgomery), z Montgomery,
 * h out: N = 2^ceil_log2(n + l) elements. */
static void fr_witness_map(const fr_ops* f, size_t n, size_t l, const uint32_t* const ptr[3],
                           const uint32_t* const col[3], const uint64_t* const val[3], const fr_t* z, fr_t* h,
                           int log_n, int threads) {
    size_t N = (size_t)1 << log_n;
    fr_t* a = (fr_t*)calloc(N, sizeof(fr_t));
    fr_t* b = (fr_t*)calloc(N, sizeof(fr_t));
    fr_t* c = (fr_t*)calloc(N, sizeof(fr_t));
    fr_t* arr[3] = {a, b, c};
    for (int m = 0; m < 3; m++) {
#pragma omp parallel for num_threads(threads) schedule(static)
        for (size_t i = 0; i < n; i++) {
            fr_t acc; memset(&acc, 0, 32);
            for (uint32_t k = ptr[m][i]; k < ptr[m][i + 1]; k++) {
                fr_t t;
                f->mul(&t, (const fr_t*)(val[m] + 4 * (size_t)k), &z[col[m][k]]);
                f->add(&acc, &acc, &t);
            }
            arr[m][i] = acc;
        }
    }
    for (size_t i = 0; i < l; i++) a[n + i] = z[i];
    for (int m = 0; m < 3; m++) {
        fr_ntt(f, arr[m], log_n, 1, 0, threads);
        fr_ntt(f, arr[m], log_n, 0, 1, threads);
    }
    fr_t g, gN, one, zi; memcpy(&g, f->gen, 32); memcpy(&one, f->one, 32);
    gN = g;
    for (int i = 0; i < log_n; i++) f->mul(&gN, &gN, &gN);
    f->sub(&gN, &gN, &one);
    f->inv(&zi, &gN);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < N; i++) {
        fr_t t;
        f->mul(&t, &a[i], &b[i]);
        f->sub(&t, &t, &c[i]);
        f->mul(&h[i], &t, &zi);
    }
    fr_ntt(f, h, log_n, 1, 1, threads);
    free(a); free(b); free(c);
}

int orc_witness_map(int curve, size_t n, size_t l, const uint32_t* a_ptr, const uint32_t* a_col, const uint64_t* a_val,
                    const uint32_t* b_ptr, const uint32_t* b_col, const uint64_t* b_val, const uint32_t* c_ptr,
                    const uint32_t* c_col, const uint64_t* c_val, const uint64_t* z, uint64_t* h, int threads) {
    fr_ops f; get_fr(curve, &f);
    int log_n = ark_log2_ceil(n + l);
    if (log_n > f.two_adicity) return -6;
    const uint32_t* ptr[3] = {a_ptr, b_ptr, c_ptr};
    const uint32_t* col[3] = {a_col, b_col, c_col};
    const uint64_t* val[3] = {a_val, b_val, c_val};
    fr_witness_map(&f, n, l, ptr, col, val, (const fr_t*)z, (fr_t*)h, log_n, threads < 1 ? 1 : threads);
    return 0;
}

/* ---- MSM entry points ----------------------------------------------------------------------------
 * bases: affine Montgomery (x||y, G2: x.c0 x.c1 y.c0 y.c1), inf: per-point flags or NULL.
 * scalars: canonical (repr 0) or Montgomery Fr (repr 1).  out: affine Montgomery + inf flag. */
#define MSM_ENTRY(NAME, G, FE_T, FRP, NUMBITS, COORD_LIMBS)                                                        \
    int NAME(const uint64_t* bases, const uint8_t* inf, const uint64_t* scalars, size_t n, int repr,               \
             uint64_t* out_xy, uint8_t* out_inf, int threads) {                                                    \
        init_all();                                                                                                \
        if (threads < 1) threads = 1;                                                                              \
        G##_aff* pts = (G##_aff*)malloc(sizeof(G##_aff) * (n ? n : 1));                                            \
        for (size_t i = 0; i < n; i++) {                                                                           \
            memcpy(&pts[i].x, bases + (size_t)i * 2 * COORD_LIMBS, 8 * COORD_LIMBS);                               \
            memcpy(&pts[i].y, bases + (size_t)i * 2 * COORD_LIMBS + COORD_LIMBS, 8 * COORD_LIMBS);                 \
            pts[i].inf = inf ? inf[i] : 0;                                                                         \
        }                                                                                                          \
        uint64_t* sc = (uint64_t*)malloc(32 * (n ? n : 1));                                                        \
        if (repr == 1) {                                                                                           \
            for (size_t i = 0; i < n; i++) FRP##_from_mont((FRP##_t*)(sc + 4 * i), (const FRP##_t*)(scalars + 4 * i)); \
        } else {                                                                                                   \
            memcpy(sc, scalars, 32 * n);                                                                           \
        }                                                                                                          \
        G##_jac r;                                                                                                 \
        G##_msm(&r, pts, sc, n, NUMBITS, threads);                                                                 \
        G##_aff a;                                                                                                 \
        G##_to_affine(&a, &r);                                                                                     \
        memcpy(out_xy, &a.x, 8 * COORD_LIMBS);                                                                     \
        memcpy(out_xy + COORD_LIMBS, &a.y, 8 * COORD_LIMBS);                                                       \
        if (out_inf) *out_inf = (uint8_t)a.inf;                                                                    \
        free(pts);                                                                                                 \
        free(sc);                                                                                                  \
        return 0;                                                                                                  \
    }
MSM_ENTRY(orc_msm_bls_g1, bls_g1, bls_fq_t, bls_fr, 255, 6)
MSM_ENTRY(orc_msm_bls_g2, bls_g2, bls_fq2_t, bls_fr, 255, 12)
MSM_ENTRY(orc_msm_bn_g1, bn_g1, bn_fq_t, bn_fr, 254, 4)
MSM_ENTRY(orc_msm_bn_g2, bn_g2, bn_fq2_t, bn_fr, 254, 8)

int orc_msm(int curve, int group, const uint64_t* bases, const uint8_t* inf, const uint64_t* scalars, size_t n,
            int repr, uint64_t* out_xy, uint8_t* out_inf, int threads) {
    if (curve == 0) return group == 1 ? orc_msm_bls_g1(bases, inf, scalars, n, repr, out_xy, out_inf, threads)
                                      : orc_msm_bls_g2(bases, inf, scalars, n, repr, out_xy, out_inf, threads);
    return group == 1 ? orc_msm_bn_g1(bases, inf, scalars, n, repr, out_xy, out_inf, threads)
                      : orc_msm_bn_g2(bases, inf, scalars, n, repr, out_xy, out_inf, threads);
}

/* element-wise Montgomery product (pins the C field code against Python ints) */
int orc_field_mul(int curve, int field, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    init_all();
    for (size_t i = 0; i < n; i++) {
        if (field == 0) {
            if (curve == 0) bls_fr_mul((bls_fr_t*)(out + 4 * i), (const bls_fr_t*)(a + 4 * i), (const bls_fr_t*)(b + 4 * i));
            else bn_fr_mul((bn_fr_t*)(out + 4 * i), (const bn_fr_t*)(a + 4 * i), (const bn_fr_t*)(b + 4 * i));
        } else {
            if (curve == 0) bls_fq_mul((bls_fq_t*)(out + 6 * i), (const bls_fq_t*)(a + 6 * i), (const bls_fq_t*)(b + 6 * i));
            else bn_fq_mul((bn_fq_t*)(out + 4 * i), (const bn_fq_t*)(a + 4 * i), (const bn_fq_t*)(b + 4 * i));
        }
    }
    return 0;
}

/* sum_i a_i * b_i mod r for canonical (non-Montgomery) 256-bit inputs (reduced mod r first); out canonical.
 * The exact expected exponent of an MSM over bases k_i * G (tests: the check shares nothing with any MSM). */
#define DOT_IMPL(NAME, FRP)                                                                                        \
    static void NAME##_red(FRP##_t* x) { while (FRP##_geq_mod(x)) FRP##_sub_mod_raw(x); }                          \
    static int NAME(const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out, int threads) {                  \
        init_all();                                                                                                \
        if (threads < 1) threads = 1;                                                                              \
        FRP##_t* part = (FRP##_t*)calloc((size_t)threads, sizeof(FRP##_t));                                        \
        _Pragma("omp parallel num_threads(threads)")                                                               \
        {                                                                                                          \
            int t = 0, nt = 1;                                                                                     \
            OMP_IDS(t, nt);                                                                                        \
            FRP##_t acc; memset(&acc, 0, sizeof acc);                                                              \
            for (size_t i = (size_t)t; i < n; i += (size_t)nt) {                                                   \
                FRP##_t x, y, p;                                                                                   \
                memcpy(&x, a + 4 * i, 32); memcpy(&y, b + 4 * i, 32);                                              \
                NAME##_red(&x); NAME##_red(&y);                                                                    \
                FRP##_to_mont(&x, &x);            /* xR */                                                         \
                FRP##_mul(&p, &x, &y);            /* xR * y / R = xy */                                            \
                FRP##_add(&acc, &acc, &p);                                                                         \
            }                                                                                                      \
            part[t] = acc;                                                                                         \
        }                                                                                                          \
        FRP##_t tot; memset(&tot, 0, sizeof tot);                                                                  \
        for (int t = 0; t < threads; t++) FRP##_add(&tot, &tot, &part[t]);                                         \
        memcpy(out, &tot, 32);                                                                                     \
        free(part);                                                                                                \
        return 0;                                                                                                  \
    }
#ifdef _OPENMP
#define OMP_IDS(t, nt) do { t = omp_get_thread_num(); nt = omp_get_num_threads(); } while (0)
#else
#define OMP_IDS(t, nt) do { } while (0)
#endif
DOT_IMPL(orc_dot_bls, bls_fr)
DOT_IMPL(orc_dot_bn, bn_fr)
int orc_dot_fr(int curve, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out, int threads) {
    return curve == 0 ? orc_dot_bls(a, b, n, out, threads) : orc_dot_bn(a, b, n, out, threads);
}

/* ---- baseline helpers ---------------------------------------------------------------------------- */
int orc_uses_asm(void) { init_all(); return g_orc_use_asm; }

/* ns per Montgomery multiplication, single thread, dependent chain; field 0 = Fr, 1 = Fq.  Printed beside the CPU
 * baseline so that the reader can compare with ark-ff's published figures. */
double orc_bench_field_mul(int curve, int field, long iters) {
    init_all();
    struct timespec t0, t1;
    volatile uint64_t sink = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (field == 1 && curve == 0) {
        bls_fq_t a = bls_fq_R2, b = bls_fq_R2; b.v[0] ^= 0x1234567;
        for (long i = 0; i < iters; i++) bls_fq_mul(&a, &a, &b);
        sink = a.v[0];
    } else if (field == 1) {
        bn_fq_t a = bn_fq_R2, b = bn_fq_R2; b.v[0] ^= 0x1234567;
        for (long i = 0; i < iters; i++) bn_fq_mul(&a, &a, &b);
        sink = a.v[0];
    } else if (curve == 0) {
        bls_fr_t a = bls_fr_R2, b = bls_fr_R2; b.v[0] ^= 0x1234567;
        for (long i = 0; i < iters; i++) bls_fr_mul(&a, &a, &b);
        sink = a.v[0];
    } else {
        bn_fr_t a = bn_fr_R2, b = bn_fr_R2; b.v[0] ^= 0x1234567;
        for (long i = 0; i < iters; i++) bn_fr_mul(&a, &a, &b);
        sink = a.v[0];
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    (void)sink;
    return ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / (double)iters;
}

/* out[i] = scalars[i] * G for an affine generator G: the fixed-base multiplications of generator.rs:128-196
 * (`FixedBase::msm`), used to build the synthetic proving key of the CPU reference arm without touching the GPU.
 * 4-bit fixed windows over a table T[w][d] = d 16^w G, 64 mixed additions per point, one batched inversion
 * (Montgomery's trick) per 256 points.  scalars canonical (n x 4 u64); out affine Montgomery, infinity as (0, 0). */
#define FIXED_BASE_IMPL(NAME, G, FE, COORD_LIMBS)                                                                  \
    static int NAME(const uint64_t* gen_xy, const uint64_t* scalars, size_t n, uint64_t* out_xy, int threads) {     \
        init_all();                                                                                                \
        if (threads < 1) threads = 1;                                                                              \
        G##_aff g;                                                                                                 \
        memcpy(&g.x, gen_xy, 8 * COORD_LIMBS); memcpy(&g.y, gen_xy + COORD_LIMBS, 8 * COORD_LIMBS); g.inf = 0;     \
        G##_aff* tbl = (G##_aff*)malloc(sizeof(G##_aff) * 64 * 16);                                                \
        G##_jac base; base.x = g.x; base.y = g.y; base.z = FE##_ONE;                                               \
        for (int w = 0; w < 64; w++) {                                                                             \
            G##_jac acc; G##_set_inf(&acc);                                                                        \
            G##_aff b_aff; G##_to_affine(&b_aff, &base);                                                           \
            memset(&tbl[w * 16], 0, sizeof(G##_aff)); tbl[w * 16].inf = 1;                                         \
            for (int d = 1; d < 16; d++) { G##_add_mixed(&acc, &b_aff); G##_to_affine(&tbl[w * 16 + d], &acc); }   \
            for (int k = 0; k < 4; k++) G##_dbl(&base);                                                            \
        }                                                                                                          \
        const size_t CH = 256;                                                                                     \
        const size_t nch = (n + CH - 1) / CH;                                                                      \
        _Pragma("omp parallel for num_threads(threads) schedule(dynamic, 4)")                                      \
        for (size_t c = 0; c < nch; c++) {                                                                         \
            G##_jac pt[256]; FE##_t pre[256];                                                                      \
            size_t lo = c * CH, hi = lo + CH < n ? lo + CH : n;                                                    \
            FE##_t run = FE##_ONE;                                                                                 \
            for (size_t i = lo; i < hi; i++) {                                                                     \
                G##_jac acc; G##_set_inf(&acc);                                                                    \
                const uint64_t* k = scalars + 4 * i;                                                               \
                for (int w = 0; w < 64; w++) {                                                                     \
                    unsigned d = (unsigned)(k[w >> 4] >> ((w & 15) * 4)) & 15u;                                    \
                    if (d) G##_add_mixed(&acc, &tbl[w * 16 + d]);                                                  \
                }                                                                                                  \
                pt[i - lo] = acc;                                                                                  \
                pre[i - lo] = run;                                                                                 \
                if (!G##_is_inf(&acc)) FE##_mul(&run, &run, &acc.z);                                               \
            }                                                                                                      \
            FE##_t inv; FE##_inv(&inv, &run);                                                                      \
            for (size_t i = hi; i-- > lo;) {                                                                       \
                uint64_t* o = out_xy + i * 2 * COORD_LIMBS;                                                        \
                G##_jac* p = &pt[i - lo];                                                                          \
                if (G##_is_inf(p)) { memset(o, 0, 16 * COORD_LIMBS); continue; }                                   \
                FE##_t zi, zi2, t;                                                                                 \
                FE##_mul(&zi, &inv, &pre[i - lo]);                                                                 \
                FE##_mul(&inv, &inv, &p->z);                                                                       \
                FE##_sqr(&zi2, &zi);                                                                               \
                FE##_mul(&t, &p->x, &zi2); memcpy(o, &t, 8 * COORD_LIMBS);                                         \
                FE##_mul(&zi2, &zi2, &zi);                                                                         \
                FE##_mul(&t, &p->y, &zi2); memcpy(o + COORD_LIMBS, &t, 8 * COORD_LIMBS);                           \
            }                                                                                                      \
        }                                                                                                          \
        free(tbl);                                                                                                 \
        return 0;                                                                                                  \
    }
FIXED_BASE_IMPL(orc_fixed_bls_g1, bls_g1, bls_fq, 6)
FIXED_BASE_IMPL(orc_fixed_bls_g2, bls_g2, bls_fq2, 12)
FIXED_BASE_IMPL(orc_fixed_bn_g1, bn_g1, bn_fq, 4)
FIXED_BASE_IMPL(orc_fixed_bn_g2, bn_g2, bn_fq2, 8)
int orc_fixed_base_mul(int curve, int group, const uint64_t* gen_xy, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                       int threads) {
    if (curve == 0) return group == 1 ? orc_fixed_bls_g1(gen_xy, scalars, n, out_xy, threads)
                                      : orc_fixed_bls_g2(gen_xy, scalars, n, out_xy, threads);
    return group == 1 ? orc_fixed_bn_g1(gen_xy, scalars, n, out_xy, threads)
                      : orc_fixed_bn_g2(gen_xy, scalars, n, out_xy, threads);
}

/* ---- whole-proof restatement (prover.rs:53-156 + committer.rs:106-114) -------------------------
 * One macro per curve keeps every group operation in the reference's order. */
typedef struct {
    size_t n, l, nw;                 /* constraints, instance vars, witness vars */
    const uint32_t *a_ptr, *a_col, *b_ptr, *b_col, *c_ptr, *c_col;
    const uint64_t *a_val, *b_val, *c_val;
    uint32_t num_stages;
    const uint32_t* stage_start;     /* num_stages + 1 entries */
    const uint64_t *a_g, *b_g, *b_h, *h_g;       /* m, m, m, N-1 affine points */
    const uint64_t* const* ck;       /* per-stage committer keys */
    const uint64_t* deltas_g;        /* num_stages points */
    const uint64_t *alpha_g, *beta_g, *beta_h, *last_delta_h;
} orc_pk;

#define PROVE_IMPL(NAME, C, G1, G2, FRP, FQL, NUMBITS)                                                             \
    static void NAME##_load1(G1##_aff* p, const uint64_t* src) {                                                   \
        memcpy(&p->x, src, 8 * FQL); memcpy(&p->y, src + FQL, 8 * FQL);                                            \
        p->inf = G1##_aff_is_zero_xy(p);                                                                           \
    }                                                                                                              \
    int NAME(const orc_pk* pk, const uint64_t* z_mont, const uint64_t* r_mont, const uint64_t* s_mont,             \
             const uint64_t* kappas_mont, uint64_t* outA, uint64_t* outB, uint64_t* outC, int threads) {           \
        init_all();                                                                                                \
        if (threads < 1) threads = 1;                                                                              \
        fr_ops f; get_fr(C, &f);                                                                                   \
        const size_t m = pk->l + pk->nw;                                                                           \
        const int log_n = ark_log2_ceil(pk->n + pk->l);                                                            \
        const size_t N = (size_t)1 << log_n;                                                                       \
        /* prover.rs:78-82: assignment = full_assignment[1..].into_bigint() */                                     \
        uint64_t* asg = (uint64_t*)malloc(32 * m);                                                                 \
        for (size_t i = 1; i < m; i++) FRP##_from_mont((FRP##_t*)(asg + 4 * (i - 1)), (const FRP##_t*)(z_mont + 4 * i)); \
        uint64_t rr[4], ss[4], rs[4];                                                                              \
        FRP##_from_mont((FRP##_t*)rr, (const FRP##_t*)r_mont);                                                     \
        FRP##_from_mont((FRP##_t*)ss, (const FRP##_t*)s_mont);                                                     \
        { FRP##_t t; FRP##_mul(&t, (const FRP##_t*)r_mont, (const FRP##_t*)s_mont); FRP##_from_mont((FRP##_t*)rs, &t); } \
        G1##_aff* q1 = (G1##_aff*)malloc(sizeof(G1##_aff) * (m > N ? m : N));                                      \
        G2##_aff* q2 = (G2##_aff*)malloc(sizeof(G2##_aff) * m);                                                    \
        G1##_aff delta_g, alpha_g, beta_g;                                                                         \
        NAME##_load1(&delta_g, pk->deltas_g + (size_t)(pk->num_stages - 1) * 2 * FQL);                             \
        NAME##_load1(&alpha_g, pk->alpha_g);                                                                       \
        NAME##_load1(&beta_g, pk->beta_g);                                                                         \
        /* A = r*delta + a_g[0] + MSM(a_g[1..], asg) + alpha   (prover.rs:86-88,158-171) */                        \
        for (size_t i = 0; i < m; i++) NAME##_load1(&q1[i], pk->a_g + i * 2 * FQL);                                \
        G1##_jac A, acc;                                                                                           \
        G1##_mul_u256(&A, &delta_g, rr);                                                                           \
        G1##_add_mixed(&A, &q1[0]);                                                                                \
        G1##_msm(&acc, q1 + 1, asg, m - 1, NUMBITS, threads);                                                      \
        G1##_add(&A, &acc);                                                                                        \
        G1##_add_mixed(&A, &alpha_g);                                                                              \
        /* B in G1 (prover.rs:92-102; skipped when r == 0) */                                                      \
        G1##_jac B1; G1##_set_inf(&B1);                                                                            \
        if (rr[0] | rr[1] | rr[2] | rr[3]) {                                                                       \
            for (size_t i = 0; i < m; i++) NAME##_load1(&q1[i], pk->b_g + i * 2 * FQL);                            \
            G1##_mul_u256(&B1, &delta_g, ss);                                                                      \
            G1##_add_mixed(&B1, &q1[0]);                                                                           \
            G1##_msm(&acc, q1 + 1, asg, m - 1, NUMBITS, threads);                                                  \
            G1##_add(&B1, &acc);                                                                                   \
            G1##_add_mixed(&B1, &beta_g);                                                                          \
        }                                                                                                          \
        /* B in G2 (prover.rs:106-107) */                                                                          \
        G2##_aff delta_h, beta_h;                                                                                  \
        memcpy(&delta_h.x, pk->last_delta_h, 16 * FQL); memcpy(&delta_h.y, pk->last_delta_h + 2 * FQL, 16 * FQL); delta_h.inf = 0; \
        memcpy(&beta_h.x, pk->beta_h, 16 * FQL); memcpy(&beta_h.y, pk->beta_h + 2 * FQL, 16 * FQL); beta_h.inf = 0; \
        for (size_t i = 0; i < m; i++) {                                                                           \
            memcpy(&q2[i].x, pk->b_h + i * 4 * FQL, 16 * FQL);                                                     \
            memcpy(&q2[i].y, pk->b_h + i * 4 * FQL + 2 * FQL, 16 * FQL);                                           \
            q2[i].inf = G2##_aff_is_zero_xy(&q2[i]);                                                               \
        }                                                                                                          \
        G2##_jac B2, acc2;                                                                                         \
        G2##_mul_u256(&B2, &delta_h, ss);                                                                          \
        G2##_add_mixed(&B2, &q2[0]);                                                                               \
        G2##_msm(&acc2, q2 + 1, asg, m - 1, NUMBITS, threads);                                                     \
        G2##_add(&B2, &acc2);                                                                                      \
        G2##_add_mixed(&B2, &beta_h);                                                                              \
        /* L (prover.rs:111-117): E::G1::msm(last_ck, current_witness) -- Montgomery scalars */                    \
        const uint32_t ns = pk->num_stages;                                                                        \
        const size_t nl = pk->stage_start[ns] - pk->stage_start[ns - 1];                                           \
        uint64_t* wl = (uint64_t*)malloc(32 * (nl ? nl : 1));                                                      \
        for (size_t i = 0; i < nl; i++)                                                                            \
            FRP##_from_mont((FRP##_t*)(wl + 4 * i), (const FRP##_t*)(z_mont + 4 * (pk->l + pk->stage_start[ns - 1] + i))); \
        for (size_t i = 0; i < nl; i++) NAME##_load1(&q1[i], pk->ck[ns - 1] + i * 2 * FQL);                        \
        G1##_jac Lacc;                                                                                             \
        G1##_msm(&Lacc, q1, wl, nl, NUMBITS, threads);                                                             \
        /* h = witness_map (prover.rs:123); H = MSM(h_g, h[..N-1]) (prover.rs:128-129) */                          \
        fr_t* h = (fr_t*)malloc(sizeof(fr_t) * N);                                                                 \
        { const uint32_t* ptr[3] = {pk->a_ptr, pk->b_ptr, pk->c_ptr};                                              \
          const uint32_t* col[3] = {pk->a_col, pk->b_col, pk->c_col};                                              \
          const uint64_t* val[3] = {pk->a_val, pk->b_val, pk->c_val};                                              \
          fr_witness_map(&f, pk->n, pk->l, ptr, col, val, (const fr_t*)z_mont, h, log_n, threads); }               \
        for (size_t i = 0; i + 1 < N; i++) FRP##_from_mont((FRP##_t*)&h[i], (const FRP##_t*)&h[i]);                \
        for (size_t i = 0; i + 1 < N; i++) NAME##_load1(&q1[i], pk->h_g + i * 2 * FQL);                            \
        G1##_jac Hacc;                                                                                             \
        G1##_msm(&Hacc, q1, (const uint64_t*)h, N - 1, NUMBITS, threads);                                          \
        /* C = s*A + r*B1 - rs*delta + L + H   (prover.rs:135-147) */                                              \
        G1##_aff Aaff, B1aff;                                                                                      \
        G1##_to_affine(&Aaff, &A);                                                                                 \
        G1##_to_affine(&B1aff, &B1);                                                                               \
        G1##_jac Cc, t;                                                                                            \
        G1##_mul_u256(&Cc, &Aaff, ss);                                                                             \
        G1##_mul_u256(&t, &B1aff, rr);                                                                             \
        G1##_add(&Cc, &t);                                                                                         \
        G1##_mul_u256(&t, &delta_g, rs);                                                                           \
        NAME##_neg_jac(&t);                                                                                        \
        G1##_add(&Cc, &t);                                                                                         \
        G1##_add(&Cc, &Lacc);                                                                                      \
        G1##_add(&Cc, &Hacc);                                                                                      \
        /* committer.rs:112-114: C -= msm_unchecked(deltas_g, comm_rands) (zips to the shorter) */                 \
        for (uint32_t i = 0; i + 1 < ns; i++) {                                                                    \
            uint64_t k[4];                                                                                         \
            FRP##_from_mont((FRP##_t*)k, (const FRP##_t*)(kappas_mont + 4 * i));                                   \
            G1##_aff d; NAME##_load1(&d, pk->deltas_g + (size_t)i * 2 * FQL);                                      \
            G1##_mul_u256(&t, &d, k);                                                                              \
            NAME##_neg_jac(&t);                                                                                    \
            G1##_add(&Cc, &t);                                                                                     \
        }                                                                                                          \
        G1##_aff Caff; G2##_aff Baff;                                                                              \
        G1##_to_affine(&Caff, &Cc);                                                                                \
        G2##_to_affine(&Baff, &B2);                                                                                \
        memcpy(outA, &Aaff.x, 8 * FQL); memcpy(outA + FQL, &Aaff.y, 8 * FQL);                                      \
        memcpy(outB, &Baff.x, 16 * FQL); memcpy(outB + 2 * FQL, &Baff.y, 16 * FQL);                                \
        memcpy(outC, &Caff.x, 8 * FQL); memcpy(outC + FQL, &Caff.y, 8 * FQL);                                      \
        free(asg); free(q1); free(q2); free(wl); free(h);                                                          \
        return 0;                                                                                                  \
    }

#define AFF_HELPERS(G1, G2, FQ, FQ2, NAME)                                                                         \
    static int G1##_aff_is_zero_xy(const G1##_aff* p) { return FQ##_is_zero(&p->x) && FQ##_is_zero(&p->y); }       \
    static int G2##_aff_is_zero_xy(const G2##_aff* p) { return FQ2##_is_zero(&p->x) && FQ2##_is_zero(&p->y); }     \
    static void NAME##_neg_jac(G1##_jac* p) { FQ##_neg(&p->y, &p->y); }
AFF_HELPERS(bls_g1, bls_g2, bls_fq, bls_fq2, orc_prove_bls)
AFF_HELPERS(bn_g1, bn_g2, bn_fq, bn_fq2, orc_prove_bn)
PROVE_IMPL(orc_prove_bls, 0, bls_g1, bls_g2, bls_fr, 6, 255)
PROVE_IMPL(orc_prove_bn, 1, bn_g1, bn_g2, bn_fr, 4, 254)

int orc_prove(int curve, const orc_pk* pk, const uint64_t* z, const uint64_t* r, const uint64_t* s,
              const uint64_t* kappas, uint64_t* outA, uint64_t* outB, uint64_t* outC, int threads) {
    return curve == 0 ? orc_prove_bls(pk, z, r, s, kappas, outA, outB, outC, threads)
                      : orc_prove_bn(pk, z, r, s, kappas, outA, outB, outC, threads);
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
