// Device-resident constraint matrices + witness-map entry points (see qap.cu).
#pragma once
#include "common.cuh"

namespace hk {

struct R1csDev {
    size_t n = 0, l = 0, nw = 0;   // constraints, instance variables (incl. one), witness variables
    uint32_t log_n = 0;            // domain size = next_pow2(n + l)  (generator.rs:66-67)
    size_t nnz[3] = {0, 0, 0};
    DevBuf b_ptr[3], b_col[3], b_val[3];
    uint32_t* ptr[3] = {nullptr, nullptr, nullptr};
    uint32_t* col[3] = {nullptr, nullptr, nullptr};
    uint64_t* val[3] = {nullptr, nullptr, nullptr};
};

std::shared_ptr<R1csDev> r1cs_upload(hk_ctx* ctx, size_t n, size_t l, size_t nw, const uint32_t* const ptr[3],
                                     const uint32_t* const col[3], const uint64_t* const val[3]);

// abc: device buffer of K x 3N Fr (a | b | c per proof), clobbered; h: device buffer of K x N Fr.
template <class Curve>
void qap_h_from_evals(hk_ctx* ctx, uint32_t log_n, uint64_t* d_abc, uint64_t* d_h, size_t K = 1);

// per proof k: abc[k][0..N) = a evals, [N..2N) = b, [2N..3N) = c from the CSR matrices and z[k].
template <class Curve>
void r1cs_evals(hk_ctx* ctx, const R1csDev& m, const uint64_t* d_z, uint64_t* d_abc, size_t K = 1);

}  // namespace hk
