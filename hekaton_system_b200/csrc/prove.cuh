// Proving-key residency + fused commit/prove entry points (see prove.cu).
#pragma once
#include "common.cuh"

namespace hk {

struct R1csDev;

struct PkDev {
    std::shared_ptr<R1csDev> r1cs;
    uint32_t num_stages = 0;
    size_t m = 0;                                   // variables incl. the constant one
    std::vector<size_t> stage_start;                // witness-index ranges per stage
    BaseSet a_ext, b1_ext, b2_ext;                  // [q[1..], q[0], vk_param, delta]  (+ window tables)
    std::vector<std::unique_ptr<BaseSet>> ck_ext;   // per stage: [ck.., delta_last (, deltas_g[0..ns-2])]
    std::shared_ptr<BaseSet> h_g;                   // shared with the caller's h_g base set
    uint64_t h_g_handle = 0;
    bool share_b_sort = false;                      // b1_ext / b2_ext agree in geometry and infinity pattern
};

template <class Curve>
void apply_infinity(hk_ctx* ctx, BaseSet& bs, const uint8_t* d_inf);
template <class Curve, int GROUP>
void jac_to_affine(hk_ctx* ctx, const void* d_jac, void* d_aff, uint8_t* d_inf, size_t n, cudaStream_t st = nullptr);
template <class Curve, int GROUP>
void fixed_base_mul(hk_ctx* ctx, const void* d_gen_affine, const Scalar256* d_scalars, size_t n, bool mont,
                    void* d_out_affine);
// out[i] = scalars[i] * points[i] (affine in, affine out): pairing_ops.rs:32-39 `scalar_pairing`
template <class Curve, int GROUP>
void scalar_mul_batch(hk_ctx* ctx, const void* d_pts, const uint8_t* d_inf, const Scalar256* d_scalars, size_t n, bool mont,
                      void* d_out_affine, uint8_t* d_inf_out);
template <class Curve>
std::shared_ptr<PkDev> pk_create(hk_ctx* ctx, const hk_pk_desc& d);
template <class Curve>
void commit_run(hk_ctx* ctx, PkDev& pk, uint32_t stage, size_t K, const uint64_t* d_z, size_t z_stride,
                const uint64_t* d_kappa, void* d_out_affine, uint8_t* d_inf);
// K proofs against the same key in one pipeline: z K x m, r/s K, kappas K x (ns-1); outputs K each
template <class Curve>
void prove_run(hk_ctx* ctx, PkDev& pk, size_t K, const uint64_t* d_z, const uint64_t* d_r, const uint64_t* d_s,
               const uint64_t* d_kappas, void* d_outA, void* d_outB, void* d_outC, uint8_t* d_inf);

}  // namespace hk
