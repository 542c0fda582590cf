// Short-Weierstrass (a = 0) group arithmetic for G1 (F = Fq) and G2 (F = Fq2).
//
// Accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity ZZ = 0):
// a mixed add of an affine base costs 8M + 2S, a full add 12M + 2S, a doubling 6M + 4S
// (EFD "madd-2008-s", "add-2008-s", "dbl-2008-s-1").  Results cross the C ABI as
// Jacobian (X, Y, Z) because that is what ark-ec 0.4 `short_weierstrass::Projective`
// holds; any representative of the group element is acceptable there,
// equality is defined after `into_affine` (prover.rs:152-154).
//
// An affine base at infinity is encoded as (0, 0), which is not on any of the four
// curves (b != 0).
#pragma once
#include "field.cuh"

namespace hk {

template <class F>
struct Affine {
    F x, y;
    __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
    __device__ __forceinline__ static Affine inf() { return {F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    __device__ __forceinline__ static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }
    __device__ __forceinline__ static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, F::one(), F::one()};
    }
    __device__ __forceinline__ XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

    // this = 2 * (affine p), p finite with y != 0 handled by caller via is_inf of result
    __device__ __forceinline__ static XYZZ dbl_affine(const Affine<F>& p) {
        // mdbl-2008-s-1: U = 2Y, V = U^2, W = U V, S = X V, M = 3 X^2
        F U = p.y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = p.x * V;
        F x2 = p.x.sqr();
        F M = x2.dbl() + x2;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * p.y;
        return {X3, Y3, V, W};
    }

    __device__ __forceinline__ XYZZ dbl() const {
        if (is_inf()) return *this;
        // dbl-2008-s-1 with a = 0.  (y = 0 cannot occur: no 2-torsion in odd-order groups;
        // if it did, V = 0 and the result is the identity, which is also correct.)
        F U = y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = x * V;
        F x2 = x.sqr();
        F M = x2.dbl() + x2;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * y;
        return {X3, Y3, V * zz, W * zzz};
    }

    // this += affine p   (negate selects this -= p)
    __device__ __forceinline__ void madd(const Affine<F>& p_in, bool negate = false) {
        if (p_in.is_inf()) return;
        Affine<F> p = p_in;
        if (negate) p.y = p.y.neg();
        if (is_inf()) {
            *this = {p.x, p.y, F::one(), F::one()};
            return;
        }
        F U2 = p.x * zz;
        F S2 = p.y * zzz;
        F P = U2 - x;
        F R = S2 - y;
        if (P.is_zero()) {
            if (R.is_zero()) {
                *this = dbl_affine(p);
            } else {
                *this = inf();
            }
            return;
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Q = x * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        F Y3 = F::mul_sub_mul(R, Q - X3, y, PPP);
        x = X3;
        y = Y3;
        zz = zz * PP;
        zzz = zzz * PPP;
    }

    // this += q
    __device__ __forceinline__ void add(const XYZZ& q) {
        if (q.is_inf()) return;
        if (is_inf()) {
            *this = q;
            return;
        }
        F U1 = x * q.zz;
        F U2 = q.x * zz;
        F S1 = y * q.zzz;
        F S2 = q.y * zzz;
        F P = U2 - U1;
        F R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) {
                *this = dbl();
            } else {
                *this = inf();
            }
            return;
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        F Y3 = F::mul_sub_mul(R, Q - X3, S1, PPP);
        x = X3;
        y = Y3;
        zz = zz * q.zz * PP;
        zzz = zzz * q.zzz * PPP;
    }

    // k * this for a small unsigned k (double-and-add, MSB first)
    __device__ __forceinline__ XYZZ mul_small(uint32_t k) const {
        XYZZ acc = inf();
        for (int b = 31 - __clz(k | 1); b >= 0; b--) {
            acc = acc.dbl();
            if ((k >> b) & 1) acc.add(*this);
        }
        if (k == 0) return inf();
        return acc;
    }

    // affine normalisation (one field inversion); identity -> (0, 0)
    __device__ __forceinline__ Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        // 1/ZZ and 1/ZZZ from one inversion: i = 1/(ZZ*ZZZ)
        F i = (zz * zzz).inverse_fast();
        F izz = i * zzz;
        F izzz = i * zz;
        return {x * izz, y * izzz};
    }
};

// ---------------------------------------------------------------------------------------------------------
// Quad-cooperative point arithmetic for the LATENCY-bound kernels (partial-list folds, bucket reduction, tree
// sums, the final scalar multiplications).  A warp issues one IMAD.WIDE per 4 cycles whatever its active lanes, so a
// point addition executed by one lane costs the issue time of its 14 dependent-or-not multiplications in sequence
// (~9 us per addition per warp), and a kernel whose threads each walk a chain of additions is bound by that.
// Here FOUR adjacent lanes (a quad) hold replicated copies of the operands and each computes one of the independent
// products of a round; results are exchanged with width-4 shuffles.  add: 4 rounds instead of 14 products in
// sequence, dbl: 3 instead of 10.  Throughput per warp is about the same (8 additions in 4 product-times against 32
// in 14), so the quad kernels replace the single-lane ones wherever a grid cannot fill the GPU anyway.
// All predicates are evaluated on replicated values, so every branch is uniform within a quad and the quad-masked
// shuffles are always converged.  blockDim.x must be a multiple of 4.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned quad_lane() { return threadIdx.x & 3u; }
__device__ __forceinline__ unsigned quad_mask() { return 0xFu << ((threadIdx.x & 31u) & ~3u); }

template <class Tag>
__device__ __forceinline__ Fp<Tag> quad_bcast(const Fp<Tag>& v, int src) {
    Fp<Tag> o;
    const unsigned m = quad_mask();
#pragma unroll
    for (int i = 0; i < Fp<Tag>::N64; i++) o.r.v[i] = __shfl_sync(m, v.r.v[i], src, 4);
    return o;
}
template <class F>
__device__ __forceinline__ Fp2T<F> quad_bcast(const Fp2T<F>& v, int src) {
    return {quad_bcast(v.c0, src), quad_bcast(v.c1, src)};
}
template <class Tag>
__device__ __forceinline__ Fp<Tag> quad_pick(unsigned q, const Fp<Tag>& a0, const Fp<Tag>& a1, const Fp<Tag>& a2,
                                             const Fp<Tag>& a3) {
    Fp<Tag> o;
#pragma unroll
    for (int i = 0; i < Fp<Tag>::N64; i++) {
        const uint64_t lo = (q & 1u) ? a1.r.v[i] : a0.r.v[i];
        const uint64_t hi = (q & 1u) ? a3.r.v[i] : a2.r.v[i];
        o.r.v[i] = (q & 2u) ? hi : lo;
    }
    return o;
}
template <class F>
__device__ __forceinline__ Fp2T<F> quad_pick(unsigned q, const Fp2T<F>& a0, const Fp2T<F>& a1, const Fp2T<F>& a2,
                                             const Fp2T<F>& a3) {
    return {quad_pick(q, a0.c0, a1.c0, a2.c0, a3.c0), quad_pick(q, a0.c1, a1.c1, a2.c1, a3.c1)};
}

// 2 * p, operands and result replicated over the quad (dbl-2008-s-1, a = 0)
template <class F>
__device__ __noinline__ XYZZ<F> quad_dbl(const XYZZ<F>& p) {
    if (p.is_inf()) return p;
    const unsigned q = quad_lane();
    const F U = p.y.dbl();
    // round 1: V = U^2 | x2 = x^2
    F l = quad_pick(q, U, p.x, U, p.x);
    const F m1 = l * l;
    const F V = quad_bcast(m1, 0), x2 = quad_bcast(m1, 1);
    const F M = x2.dbl() + x2;
    // round 2: W = U V | S = x V | M^2 | ZZ3 = V zz
    l = quad_pick(q, U, p.x, M, V);
    F r = quad_pick(q, V, V, M, p.zz);
    const F m2 = l * r;
    const F W = quad_bcast(m2, 0), S = quad_bcast(m2, 1), MM = quad_bcast(m2, 2), ZZ3 = quad_bcast(m2, 3);
    const F X3 = MM - S.dbl();
    // round 3: M (S - X3) | W y | ZZZ3 = W zzz
    const F SX = S - X3;
    l = quad_pick(q, M, W, W, M);
    r = quad_pick(q, SX, p.y, p.zzz, SX);
    const F m3 = l * r;
    const F T1 = quad_bcast(m3, 0), T2 = quad_bcast(m3, 1), ZZZ3 = quad_bcast(m3, 2);
    return {X3, T1 - T2, ZZ3, ZZZ3};
}

// a + b, operands and result replicated over the quad (add-2008-s)
template <class F>
__device__ __noinline__ XYZZ<F> quad_add(const XYZZ<F>& a, const XYZZ<F>& b) {
    if (b.is_inf()) return a;
    if (a.is_inf()) return b;
    const unsigned q = quad_lane();
    // round 1: U1 = x1 zz2 | U2 = x2 zz1 | S1 = y1 zzz2 | S2 = y2 zzz1
    F l = quad_pick(q, a.x, b.x, a.y, b.y);
    F r = quad_pick(q, b.zz, a.zz, b.zzz, a.zzz);
    const F m1 = l * r;
    const F U1 = quad_bcast(m1, 0), U2 = quad_bcast(m1, 1), S1 = quad_bcast(m1, 2), S2 = quad_bcast(m1, 3);
    const F P = U2 - U1, R = S2 - S1;
    if (P.is_zero()) {
        if (R.is_zero()) return quad_dbl(a);
        return XYZZ<F>::inf();
    }
    // round 2: PP = P^2 | RR = R^2 | zz1 zz2 | zzz1 zzz2
    l = quad_pick(q, P, R, a.zz, a.zzz);
    r = quad_pick(q, P, R, b.zz, b.zzz);
    const F m2 = l * r;
    const F PP = quad_bcast(m2, 0), RR = quad_bcast(m2, 1), ZZm = quad_bcast(m2, 2), ZZZm = quad_bcast(m2, 3);
    // round 3: PPP = P PP | Q = U1 PP | ZZ3 = (zz1 zz2) PP
    l = quad_pick(q, P, U1, ZZm, P);
    const F m3 = l * PP;
    const F PPP = quad_bcast(m3, 0), Q = quad_bcast(m3, 1), ZZ3 = quad_bcast(m3, 2);
    const F X3 = RR - PPP - Q.dbl();
    // round 4: R (Q - X3) | S1 PPP | ZZZ3 = (zzz1 zzz2) PPP
    const F QX = Q - X3;
    l = quad_pick(q, R, S1, ZZZm, R);
    r = quad_pick(q, QX, PPP, PPP, QX);
    const F m4 = l * r;
    const F T1 = quad_bcast(m4, 0), T2 = quad_bcast(m4, 1), ZZZ3 = quad_bcast(m4, 2);
    return {X3, T1 - T2, ZZ3, ZZZ3};
}

// k * p for a small unsigned k (double-and-add, MSB first), quad-cooperative
template <class F>
__device__ __forceinline__ XYZZ<F> quad_mul_small(const XYZZ<F>& p, uint32_t k) {
    if (k == 0) return XYZZ<F>::inf();
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int b = 31 - __clz(k); b >= 0; b--) {
        acc = quad_dbl(acc);
        if ((k >> b) & 1) acc = quad_add(acc, p);
    }
    return acc;
}

// Jacobian triple as ark-ec Projective holds it.
template <class F>
struct Jacobian {
    F x, y, z;
    __device__ __forceinline__ static Jacobian from_xyzz(const XYZZ<F>& p) {
        if (p.is_inf()) return {F::one(), F::one(), F::zero()};
        // Z = ZZ  =>  X' = x Z^2 = X*ZZ,  Y' = y Z^3 = Y*ZZ^3/ZZZ = Y*ZZZ
        return {p.x * p.zz, p.y * p.zzz, p.zz};
    }
    __device__ __forceinline__ XYZZ<F> to_xyzz() const {
        if (z.is_zero()) return XYZZ<F>::inf();
        F zz = z.sqr();
        return {x, y, zz, zz * z};
    }
};

}  // namespace hk
