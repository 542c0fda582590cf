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
