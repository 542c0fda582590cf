// Explicit instantiation of the MSM pipeline for Bls12381 G2 (one TU per group keeps builds parallel).
#include "msm_impl.cuh"
namespace hk {
template void msm_run<Bls12381, 2>(hk_ctx*, const MsmBases&, size_t, size_t, const Scalar256*, size_t, bool, void*, size_t, int);
template void bases_precompute<Bls12381, 2>(hk_ctx*, BaseSet&, double);
}
