// Explicit instantiation of the MSM pipeline for Bn254 G1 (one TU per group keeps builds parallel).
// One copy of the Montgomery multiply per kernel instead of one per call site: measured +12 % on the
// proof pipeline and -15 % on the G2 MSM (instruction-cache pressure).
#define HK_MUL_CALL 1
#include "msm_impl.cuh"
namespace hk {
template void msm_run<Bn254, 1>(hk_ctx*, const MsmBases&, size_t, size_t, const Scalar256*, size_t, bool, void*, size_t, int);
template void bases_precompute<Bn254, 1>(hk_ctx*, BaseSet&, double);
}
