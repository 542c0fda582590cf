// Explicit instantiation of the MSM pipeline for Bls12381 G1 (one TU per group keeps builds parallel).
// One copy of the Montgomery multiply per kernel instead of one per call site: measured +12 % on the
// proof pipeline and -15 % on the G2 MSM (instruction-cache pressure).
#define HK_MUL_CALL 1
#ifndef HK_NO_MUL2
#define HK_MUL2 1   // a*b - c*d as one fused dual product (field.cuh mul_sub_mul): -6 % on the bucket accumulation
#endif
#ifndef HK_NO_FP2_DUAL
#define HK_FP2_DUAL 1   // Fq2 multiply = two fused dual products instead of Karatsuba's 3 multiplies + 5 additions
#endif
#ifndef HK_NO_FP2_FUSED_CALL
#define HK_FP2_FUSED_CALL 1   // whole Fq2 multiply / square as one call (field.cuh fp2_mul_call)
#endif
#include "msm_impl.cuh"
namespace hk {
template void msm_run<Bls12381, 1>(hk_ctx*, const MsmBases&, size_t, size_t, const Scalar256*, size_t, bool, void*, size_t, int);
template void bases_precompute<Bls12381, 1>(hk_ctx*, BaseSet&, double);
template void msm_accumulate<Bls12381, 1>(hk_ctx*, const MsmSorted&, const MsmBases&, void*, int);
template MsmSorted msm_sort<Bls12381>(hk_ctx*, const MsmBases&, size_t, size_t, const Scalar256*, size_t, bool, size_t, int);
}

// Which arithmetic variants this build's MSM kernels use (bench.py reports the executed multiply count from it).
extern "C" __attribute__((visibility("default"))) uint32_t hk_build_flags(void) {
    uint32_t f = 0;
#ifdef HK_MUL2
    f |= 1u;      // HK_BUILD_MUL2
#endif
#ifdef HK_MUL_KARA
    f |= 2u;      // HK_BUILD_MUL_KARA
#endif
#if defined(HK_FP2_DUAL) && defined(HK_MUL2)
    f |= 4u;      // HK_BUILD_FP2_DUAL
#endif
    return f;
}
