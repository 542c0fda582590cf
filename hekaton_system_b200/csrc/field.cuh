// Field wrappers over the generated Montgomery kernels (field_gen.cuh).
//
// Values are kept in Montgomery form, little-endian 64-bit lanes -- the same memory
// layout as ark-ff 0.4 `Fp<MontBackend<_, N>, N>` (`BigInt<N>([u64; N])`), which is what
// crosses the drop-in boundary.  Device only.
#pragma once
#include <cstdint>
#include "field_gen.cuh"

namespace hk {

// Tag types binding a generated field to the template below.
#define HK_DEFINE_FIELD_TAG(NAME)                                                             \
    struct NAME##_tag {                                                                       \
        using raw_t = NAME##_t;                                                               \
        using params = NAME##_params;                                                         \
        static constexpr int N64 = NAME##_params::N64;                                        \
        static constexpr int BITS = NAME##_params::BITS;                                      \
        __device__ __forceinline__ static void mul(raw_t& r, const raw_t& a, const raw_t& b) { NAME##_mul(r, a, b); } \
        __device__ __forceinline__ static void mulk(raw_t& r, const raw_t& a, const raw_t& b) { NAME##_mulk(r, a, b); } \
        __device__ __forceinline__ static void sqr(raw_t& r, const raw_t& a) { NAME##_sqr(r, a); }            \
        static constexpr bool HAS_MUL2 = HK_HAS_MUL2_##NAME;                                  \
        __device__ __forceinline__ static void mul2(raw_t& r, const raw_t& a, const raw_t& b, const raw_t& c, const raw_t& d) { NAME##_mul2(r, a, b, c, d); } \
        __device__ __forceinline__ static void negp(raw_t& r, const raw_t& a) { NAME##_negp(r, a); }          \
        __device__ __forceinline__ static void add(raw_t& r, const raw_t& a, const raw_t& b) { NAME##_add(r, a, b); } \
        __device__ __forceinline__ static void sub(raw_t& r, const raw_t& a, const raw_t& b) { NAME##_sub(r, a, b); } \
        __device__ __forceinline__ static uint64_t mod(int i) { return NAME##_MOD[i]; }       \
        __device__ __forceinline__ static uint64_t one(int i) { return NAME##_ONE[i]; }       \
        __device__ __forceinline__ static uint64_t r2(int i) { return NAME##_R2[i]; }         \
    };

HK_DEFINE_FIELD_TAG(bls12_381_fq)
HK_DEFINE_FIELD_TAG(bls12_381_fr)
HK_DEFINE_FIELD_TAG(bn254_fq)
HK_DEFINE_FIELD_TAG(bn254_fr)

// Instruction-cache option: with -DHK_MUL_CALL the Montgomery product is a real function (by-value
// arguments stay in registers: CALL.REL + ~26 MOVs), so a kernel holds ONE copy of the ~7 KB body
// instead of one per call site (a mixed addition inlines ten of them).
template <class Tag>
__device__ __noinline__ typename Tag::raw_t fp_mul_call(typename Tag::raw_t a, typename Tag::raw_t b) {
    typename Tag::raw_t r;
#ifdef HK_MUL_KARA
    Tag::mulk(r, a, b);      // Karatsuba product + separate reduction (fewer wide multiplies, more ALU glue)
#else
    Tag::mul(r, a, b);
#endif
    return r;
}

// Dedicated squaring (N(N+1)/2 + N^2 + N wide multiplies instead of 2N^2 + N), same call convention.
template <class Tag>
__device__ __noinline__ typename Tag::raw_t fp_sqr_call(typename Tag::raw_t a) {
    typename Tag::raw_t r;
    Tag::sqr(r, a);
    return r;
}

// Fused dual product (a*b + c*d with one reduction: 3N^2 + N wide multiplies instead of 4N^2 + 2N), same convention.
template <class Tag>
__device__ __noinline__ typename Tag::raw_t fp_mul2_call(typename Tag::raw_t a, typename Tag::raw_t b,
                                                         typename Tag::raw_t c, typename Tag::raw_t d) {
    typename Tag::raw_t r;
    Tag::mul2(r, a, b, c, d);
    return r;
}

// Whole Fq2 product / square in ONE call (-DHK_FP2_FUSED_CALL, default in the MSM / prove TUs):
// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u as two dual products in one body, and
// (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u as two multiplies in one body.  Half the calls (and argument moves) of
// the G2 point additions, 25 % fewer spills, and the two independent multiplies share one schedule: G2 MSM 2^16 -4 %,
// 2^20 -1.5 % (profiles/r01f_summary.md).  The 96-byte result comes back in registers.
template <class Tag>
struct RawPair {
    typename Tag::raw_t c0, c1;
};
template <class Tag>
__device__ __noinline__ RawPair<Tag> fp2_mul_call(typename Tag::raw_t a0, typename Tag::raw_t a1,
                                                  typename Tag::raw_t b0, typename Tag::raw_t b1) {
    RawPair<Tag> r;
    typename Tag::raw_t na1;
    Tag::negp(na1, a1);
    Tag::mul2(r.c0, a0, b0, na1, b1);
    Tag::mul2(r.c1, a0, b1, a1, b0);
    return r;
}
template <class Tag>
__device__ __noinline__ RawPair<Tag> fp2_sqr_call(typename Tag::raw_t a0, typename Tag::raw_t a1) {
    RawPair<Tag> r;
    typename Tag::raw_t s, d, m;
    Tag::add(s, a0, a1);
    Tag::sub(d, a0, a1);
    Tag::mul(r.c0, s, d);
    Tag::mul(m, a0, a1);
    Tag::add(r.c1, m, m);
    return r;
}

template <class Tag>
struct Fp {
    using raw_t = typename Tag::raw_t;
    static constexpr int N64 = Tag::N64;
    static constexpr int BITS = Tag::BITS;
    raw_t r;

    __device__ __forceinline__ static Fp zero() {
        Fp z;
#pragma unroll
        for (int i = 0; i < N64; i++) z.r.v[i] = 0;
        return z;
    }
    __device__ __forceinline__ static Fp one() {
        Fp z;
#pragma unroll
        for (int i = 0; i < N64; i++) z.r.v[i] = Tag::one(i);
        return z;
    }
    __device__ __forceinline__ bool is_zero() const {
        uint64_t o = 0;
#pragma unroll
        for (int i = 0; i < N64; i++) o |= r.v[i];
        return o == 0;
    }
    __device__ __forceinline__ bool operator==(const Fp& b) const {
        uint64_t o = 0;
#pragma unroll
        for (int i = 0; i < N64; i++) o |= (r.v[i] ^ b.r.v[i]);
        return o == 0;
    }
    __device__ __forceinline__ bool operator!=(const Fp& b) const { return !(*this == b); }

#ifdef HK_MUL_CALL
    __device__ __forceinline__ Fp operator*(const Fp& b) const { Fp o; o.r = fp_mul_call<Tag>(r, b.r); return o; }
#else
    __device__ __forceinline__ Fp operator*(const Fp& b) const { Fp o; Tag::mul(o.r, r, b.r); return o; }
#endif
    __device__ __forceinline__ Fp operator+(const Fp& b) const { Fp o; Tag::add(o.r, r, b.r); return o; }
    __device__ __forceinline__ Fp operator-(const Fp& b) const { Fp o; Tag::sub(o.r, r, b.r); return o; }
#ifdef HK_MUL_CALL
    __device__ __forceinline__ Fp sqr() const { Fp o; o.r = fp_sqr_call<Tag>(r); return o; }
#else
    __device__ __forceinline__ Fp sqr() const { Fp o; Tag::sqr(o.r, r); return o; }
#endif
    // Direct (inlined) entry points to the generated variants, independent of the build switches: test hooks.
    __device__ __forceinline__ Fp mul_kara(const Fp& b) const { Fp o; Tag::mulk(o.r, r, b.r); return o; }
    __device__ __forceinline__ static Fp mul_sub_mul_fused(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        Fp nc, o;
        Tag::negp(nc.r, c.r);
        Tag::mul2(o.r, a.r, b.r, nc.r, d.r);
        return o;
    }
    // a*b - c*d.  With -DHK_MUL2 (and a field with two spare bits above 3p) one fused dual product on (a, b, p - c, d).
    __device__ __forceinline__ static Fp mul_sub_mul(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#ifdef HK_MUL2
        if constexpr (Tag::HAS_MUL2) {
            Fp nc, o;
            Tag::negp(nc.r, c.r);
#ifdef HK_MUL_CALL
            o.r = fp_mul2_call<Tag>(a.r, b.r, nc.r, d.r);
#else
            Tag::mul2(o.r, a.r, b.r, nc.r, d.r);
#endif
            return o;
        } else
#endif
        return a * b - c * d;
    }
    // a*b + c*d, same switch
    __device__ __forceinline__ static Fp mul_add_mul(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#ifdef HK_MUL2
        if constexpr (Tag::HAS_MUL2) {
            Fp o;
#ifdef HK_MUL_CALL
            o.r = fp_mul2_call<Tag>(a.r, b.r, c.r, d.r);
#else
            Tag::mul2(o.r, a.r, b.r, c.r, d.r);
#endif
            return o;
        } else
#endif
        return a * b + c * d;
    }
    __device__ __forceinline__ Fp dbl() const { Fp o; Tag::add(o.r, r, r); return o; }
    static constexpr bool HAS_MUL2 = Tag::HAS_MUL2;
    using tag_t = Tag;
    // p - a, not reduced (p for a = 0): only as an operand of the fused dual product
    __device__ __forceinline__ Fp neg_raw() const { Fp o; Tag::negp(o.r, r); return o; }
    __device__ __forceinline__ Fp neg() const { Fp z = zero(); Fp o; Tag::sub(o.r, z.r, r); return o; }

    // Montgomery <-> canonical
    __device__ __forceinline__ Fp from_mont() const {
        Fp u = zero();
        u.r.v[0] = 1;
        return (*this) * u;
    }
    __device__ __forceinline__ Fp to_mont() const {
        Fp rr;
#pragma unroll
        for (int i = 0; i < N64; i++) rr.r.v[i] = Tag::r2(i);
        return (*this) * rr;
    }

    // a^(p-2); not constant time (nothing here is secret-dependent in a way that matters
    // for a prover's public outputs).  0 -> 0.
    __device__ __noinline__ Fp inverse() const {
        // exponent p - 2, scanned MSB -> LSB
        uint64_t e[N64];
#pragma unroll
        for (int i = 0; i < N64; i++) e[i] = Tag::mod(i);
        e[0] -= 2;  // p is odd and > 2: no borrow
        Fp acc = one();
        bool started = false;
        for (int i = BITS - 1; i >= 0; i--) {
            if (started) acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) {
                acc = started ? acc * (*this) : (*this);
                started = true;
            }
        }
        return acc;
    }

    // Inverse by the binary extended Euclidean algorithm: shifts, adds and compares only, so it runs on
    // the ALU pipe and leaves the multiplier pipe to other warps (Fermat costs ~570 multiplies).  Used by
    // the batch-affine MSM rounds, one inversion per thread batch.  Montgomery in, Montgomery out; 0 -> 0.
    __device__ __noinline__ Fp inverse_gcd() const {
        if (is_zero()) return zero();
        uint64_t u[N64], v[N64], x1[N64], x2[N64], p[N64];
#pragma unroll
        for (int i = 0; i < N64; i++) {
            u[i] = r.v[i];
            p[i] = Tag::mod(i);
            v[i] = p[i];
            x1[i] = 0;
            x2[i] = 0;
        }
        x1[0] = 1;
        auto is_one = [](const uint64_t* a) {
            uint64_t o = a[0] ^ 1;
#pragma unroll
            for (int i = 1; i < N64; i++) o |= a[i];
            return o == 0;
        };
        auto shr1 = [](uint64_t* a, uint64_t top) {       // a = (top:a) >> 1
#pragma unroll
            for (int i = 0; i < N64 - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
            a[N64 - 1] = (a[N64 - 1] >> 1) | (top << 63);
        };
        auto add_to = [](uint64_t* a, const uint64_t* b) -> uint64_t {      // a += b, returns carry
            unsigned __int128 c = 0;
#pragma unroll
            for (int i = 0; i < N64; i++) {
                c += (unsigned __int128)a[i] + b[i];
                a[i] = (uint64_t)c;
                c >>= 64;
            }
            return (uint64_t)c;
        };
        auto sub_from = [](uint64_t* a, const uint64_t* b) -> uint64_t {    // a -= b, returns borrow
            uint64_t bw = 0;
#pragma unroll
            for (int i = 0; i < N64; i++) {
                unsigned __int128 d = (unsigned __int128)a[i] - b[i] - bw;
                a[i] = (uint64_t)d;
                bw = (uint64_t)(d >> 64) & 1;
            }
            return bw;
        };
        auto geq = [](const uint64_t* a, const uint64_t* b) {
#pragma unroll
            for (int i = N64 - 1; i >= 0; i--) {
                if (a[i] != b[i]) return a[i] > b[i];
            }
            return true;
        };
        auto halve_mod = [&](uint64_t* x) {               // x = x / 2 mod p
            uint64_t top = 0;
            if (x[0] & 1) top = add_to(x, p);
            shr1(x, top);
        };
        auto sub_mod = [&](uint64_t* x, const uint64_t* y) {   // x = x - y mod p
            if (sub_from(x, y)) add_to(x, p);
        };
        // invariants: x1 * a = u, x2 * a = v (mod p); gcd(a, p) = 1 so the loop ends with u == 1 or v == 1
        for (int it = 0; it < 4 * BITS + 8; it++) {
            if (is_one(u) || is_one(v)) break;
            if (!(u[0] & 1)) {
                shr1(u, 0);
                halve_mod(x1);
            } else if (!(v[0] & 1)) {
                shr1(v, 0);
                halve_mod(x2);
            } else if (geq(u, v)) {
                sub_from(u, v);
                sub_mod(x1, x2);
            } else {
                sub_from(v, u);
                sub_mod(x2, x1);
            }
        }
        Fp raw;                                           // raw = (a R)^-1 = a^-1 R^-1 as an integer mod p
        const uint64_t* res = is_one(u) ? x1 : x2;
#pragma unroll
        for (int i = 0; i < N64; i++) raw.r.v[i] = res[i];
        Fp rr;
#pragma unroll
        for (int i = 0; i < N64; i++) rr.r.v[i] = Tag::r2(i);
        return (raw * rr) * rr;                           // a^-1 R^-1 * R^2 * R^-1 * R^2 * R^-1 = a^-1 R
    }

    // The inverse to use from single-thread / low-occupancy code (affine normalisation in the finish
    // kernels): ~4x fewer instructions than the Fermat ladder when the warp does not diverge.
    __device__ __forceinline__ Fp inverse_fast() const { return inverse_gcd(); }

    // small-integer constant in Montgomery form (k * R mod p) by repeated addition
    __device__ __forceinline__ static Fp from_small(unsigned k) {
        Fp acc = zero(), o = one();
        for (unsigned i = 0; i < k; i++) acc = acc + o;
        return acc;
    }
};

// Fq2 = Fq[u]/(u^2 + 1) (both BLS12-381 and BN254 use the non-residue -1).
template <class F>
struct Fp2T {
    using base = F;
    F c0, c1;
    __device__ __forceinline__ static Fp2T zero() { return {F::zero(), F::zero()}; }
    __device__ __forceinline__ static Fp2T one() { return {F::one(), F::zero()}; }
    __device__ __forceinline__ bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    __device__ __forceinline__ bool operator==(const Fp2T& b) const { return c0 == b.c0 && c1 == b.c1; }
    __device__ __forceinline__ bool operator!=(const Fp2T& b) const { return !(*this == b); }
    __device__ __forceinline__ Fp2T operator+(const Fp2T& b) const { return {c0 + b.c0, c1 + b.c1}; }
    __device__ __forceinline__ Fp2T operator-(const Fp2T& b) const { return {c0 - b.c0, c1 - b.c1}; }
    __device__ __forceinline__ Fp2T dbl() const { return {c0.dbl(), c1.dbl()}; }
    __device__ __forceinline__ Fp2T neg() const { return {c0.neg(), c1.neg()}; }
    __device__ __forceinline__ Fp2T operator*(const Fp2T& b) const {
#if defined(HK_FP2_FUSED_CALL) && defined(HK_FP2_DUAL) && defined(HK_MUL2) && defined(HK_MUL_CALL)
        if constexpr (F::HAS_MUL2) {
            RawPair<typename F::tag_t> r = fp2_mul_call<typename F::tag_t>(c0.r, c1.r, b.c0.r, b.c1.r);
            Fp2T o;
            o.c0.r = r.c0;
            o.c1.r = r.c1;
            return o;
        }
#endif
#if defined(HK_FP2_DUAL) && defined(HK_MUL2)
        // two fused dual products (2 x (3N^2 + N) wide multiplies, no additions) instead of Karatsuba's three
        // multiplies and five additions
        if constexpr (F::HAS_MUL2) {
            return {F::mul_sub_mul(c0, b.c0, c1, b.c1), F::mul_add_mul(c0, b.c1, c1, b.c0)};
        }
#endif
        F t0 = c0 * b.c0;
        F t1 = c1 * b.c1;
        F s = (c0 + c1) * (b.c0 + b.c1);
        return {t0 - t1, s - t0 - t1};
    }
    __device__ __forceinline__ static Fp2T mul_sub_mul(const Fp2T& a, const Fp2T& b, const Fp2T& c, const Fp2T& d) {
        return a * b - c * d;
    }
    // test hooks (see Fp): the base-field variants under the same Karatsuba / dual-product composition
    __device__ __forceinline__ Fp2T mul_kara(const Fp2T& b) const {
        F t0 = c0.mul_kara(b.c0);
        F t1 = c1.mul_kara(b.c1);
        F s = (c0 + c1).mul_kara(b.c0 + b.c1);
        return {t0 - t1, s - t0 - t1};
    }
    __device__ __forceinline__ static Fp2T mul_sub_mul_fused(const Fp2T& a, const Fp2T& b, const Fp2T& c, const Fp2T& d) {
        // (a*b - c*d).c0 = a0 b0 - a1 b1 - (c0 d0 - c1 d1),  .c1 = a0 b1 + a1 b0 - (c0 d1 + c1 d0)
        F r0 = F::mul_sub_mul_fused(a.c0, b.c0, a.c1, b.c1) - F::mul_sub_mul_fused(c.c0, d.c0, c.c1, d.c1);
        F r1 = F::mul_sub_mul_fused(a.c0, b.c1, a.c1.neg(), b.c0) - F::mul_sub_mul_fused(c.c0, d.c1, c.c1.neg(), d.c0);
        return {r0, r1};
    }
    __device__ __forceinline__ Fp2T sqr() const {
#if defined(HK_FP2_FUSED_CALL) && defined(HK_MUL_CALL)
        {
            RawPair<typename F::tag_t> r = fp2_sqr_call<typename F::tag_t>(c0.r, c1.r);
            Fp2T o;
            o.c0.r = r.c0;
            o.c1.r = r.c1;
            return o;
        }
#endif
        F s = c0 + c1;
        F d = c0 - c1;
        F m = c0 * c1;
        return {s * d, m.dbl()};
    }
    __device__ __forceinline__ Fp2T inverse() const {
        F n = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * n, (c1 * n).neg()};
    }
    __device__ __forceinline__ Fp2T inverse_fast() const {
        F n = (c0.sqr() + c1.sqr()).inverse_fast();
        return {c0 * n, (c1 * n).neg()};
    }
    __device__ __forceinline__ Fp2T from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
    __device__ __forceinline__ Fp2T to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
};

// One Fq2 element spread over a LANE PAIR: the even lane of a pair (threadIdx.x & 1 == 0) holds c0, the odd lane c1.
// Used by the G2 bucket accumulation (msm_impl.cuh: k_msm_accum_g2pair): an Fq2 point addition in one thread needs
// 255 registers and still spills (2 CTAs per SM, 72 % of the multiplier peak, dependent-issue bound); split over two
// lanes every lane carries half the state -- the register budget of the G1 kernel -- and issues ONE fused dual
// product per Fq2 multiplication:
//     even lane:  c0 = a0 b0 - a1 b1 = a0 b0 + (p - a1) b1          odd lane:  c1 = a1 b0 + a0 b1
// after fetching the partner's halves with two pair-wide shuffles (24 SHFL of 32 bits for BLS12-381, against 444
// wide multiplies).  Squaring: even lane (a0 + a1)(a0 - a1), odd lane (2 a1) a0 -- one multiplication each.
// Additions / subtractions are component-wise.  Predicates (is_zero, ==) combine both halves, so every branch the
// point arithmetic takes is uniform within a pair and the shuffles (masked to the pair) are always converged.
// Only for base fields with the fused dual product (BLS12-381 Fq); blockDim.x must be even.
template <class F>
struct Fp2Half {
    using base = F;
    F v;
    __device__ __forceinline__ static unsigned pair_mask() { return 3u << ((threadIdx.x & 31u) & ~1u); }
    __device__ __forceinline__ static bool odd() { return (threadIdx.x & 1u) != 0; }
    __device__ __forceinline__ static F partner(const F& x) {
        F o;
        const unsigned m = pair_mask();
#pragma unroll
        for (int i = 0; i < F::N64; i++) o.r.v[i] = __shfl_xor_sync(m, x.r.v[i], 1);
        return o;
    }
    __device__ __forceinline__ static F sel(bool take_a, const F& a, const F& b) {
        F o;
#pragma unroll
        for (int i = 0; i < F::N64; i++) o.r.v[i] = take_a ? a.r.v[i] : b.r.v[i];
        return o;
    }
    __device__ __forceinline__ static Fp2Half zero() { return {F::zero()}; }
    __device__ __forceinline__ static Fp2Half one() { return {sel(odd(), F::zero(), F::one())}; }
    __device__ __forceinline__ bool is_zero() const {
        const bool z = v.is_zero();
        const bool pz = __shfl_xor_sync(pair_mask(), z, 1) != 0;       // both lanes must reach the shuffle: no short circuit
        return z && pz;
    }
    __device__ __forceinline__ bool operator==(const Fp2Half& b) const {
        const bool e = v == b.v;
        const bool pe = __shfl_xor_sync(pair_mask(), e, 1) != 0;
        return e && pe;
    }
    __device__ __forceinline__ bool operator!=(const Fp2Half& b) const { return !(*this == b); }
    __device__ __forceinline__ Fp2Half operator+(const Fp2Half& b) const { return {v + b.v}; }
    __device__ __forceinline__ Fp2Half operator-(const Fp2Half& b) const { return {v - b.v}; }
    __device__ __forceinline__ Fp2Half dbl() const { return {v.dbl()}; }
    __device__ __forceinline__ Fp2Half neg() const { return {v.neg()}; }
    __device__ __forceinline__ Fp2Half operator*(const Fp2Half& b) const {
        static_assert(F::HAS_MUL2, "Fp2Half needs the fused dual product");
        const bool o = odd();
        const F pa = partner(v), pb = partner(b.v);
        // even: v * b.v + (p - pa) * pb        odd: v * pb + pa * b.v
        const F x2 = sel(o, pa, pa.neg_raw());
        const F y1 = sel(o, pb, b.v);
        const F y2 = sel(o, b.v, pb);
        return {F::mul_add_mul(v, y1, x2, y2)};
    }
    __device__ __forceinline__ Fp2Half sqr() const {
        const bool o = odd();
        const F pa = partner(v);
        const F x = sel(o, v.dbl(), v + pa);     // odd: 2 a1        even: a0 + a1
        const F y = sel(o, pa, v - pa);          // odd: a0          even: a0 - a1
        return {x * y};
    }
    __device__ __forceinline__ static Fp2Half mul_sub_mul(const Fp2Half& a, const Fp2Half& b, const Fp2Half& c,
                                                          const Fp2Half& d) {
        return a * b - c * d;
    }
};

}  // namespace hk
