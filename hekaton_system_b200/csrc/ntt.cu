// Explicit instantiations of the NTT for both scalar fields.
#include "ntt_impl.cuh"
namespace hk {
template void ntt_run<Bls12381>(hk_ctx*, uint64_t*, uint32_t, size_t, uint32_t, uint32_t);
template void ntt_run<Bn254>(hk_ctx*, uint64_t*, uint32_t, size_t, uint32_t, uint32_t);
}
