// R1CS -> QAP witness map on the device.  Replaces ark-groth16 0.4
// `LibsnarkReduction::witness_map_from_matrices` (re-exported at
// /root/reference/cp-groth16/src/lib.rs:9-10, called at prover.rs:123):
//   a[i] = <A_i, z>, b[i] = <B_i, z>, c[i] = <C_i, z>  (i < n);  a[n + i] = z[i] (i < l)
//   ifft(a), ifft(b), ifft(c); coset_fft(.) with offset F::GENERATOR
//   ab[i] = (a[i] b[i] - c[i]) / Z_H(g),  Z_H(g) = g^N - 1 on the whole coset
//   coset_ifft(ab)  ->  h, all N coefficients
//
// Device pipeline (no bit-reversal passes at all):
//   k_r1cs_evals            CSR sparse mat-vec, three matrices in one launch
//   DIF(w^-1)  x3 (batched) natural evaluations -> coefficients in bit-reversed positions (x N)
//   DIT(w)     x3 (batched) with pre-scale g^rev(p)/N at the load -> coset evaluations, natural
//   k_qap_pointwise         (a b - c) * 1/(g^N - 1)
//   DIF(w^-1)               -> coefficients (x N) in bit-reversed positions; the last pass
//                           multiplies by g^-j/N and stores to the natural position j
#include "ntt_impl.cuh"
#include "qap.cuh"

namespace hk {

template <class Fr>
__global__ void k_r1cs_evals(const uint32_t* __restrict__ a_ptr, const uint32_t* __restrict__ a_col,
                             const Fr* __restrict__ a_val, const uint32_t* __restrict__ b_ptr,
                             const uint32_t* __restrict__ b_col, const Fr* __restrict__ b_val,
                             const uint32_t* __restrict__ c_ptr, const uint32_t* __restrict__ c_col,
                             const Fr* __restrict__ c_val, const Fr* __restrict__ z_all, uint32_t n, uint32_t l,
                             uint32_t N, Fr* abc_all, size_t m) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const Fr* z = z_all + (size_t)blockIdx.y * m;              // blockIdx.y = proof index in the batch
    Fr* abc = abc_all + (size_t)blockIdx.y * 3 * N;
    Fr a = Fr::zero(), b = Fr::zero(), c = Fr::zero();
    if (i < n) {
        for (uint32_t k = a_ptr[i]; k < a_ptr[i + 1]; k++) a = a + a_val[k] * z[a_col[k]];
        for (uint32_t k = b_ptr[i]; k < b_ptr[i + 1]; k++) b = b + b_val[k] * z[b_col[k]];
        for (uint32_t k = c_ptr[i]; k < c_ptr[i + 1]; k++) c = c + c_val[k] * z[c_col[k]];
    } else if (i < n + l) {
        a = z[i - n];
    }
    abc[i] = a;
    abc[(size_t)N + i] = b;
    abc[2 * (size_t)N + i] = c;
}

template <class Fr>
__global__ void k_qap_pointwise(const Fr* abc_all, uint32_t N, const Fr* zh_inv, Fr* out_all) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const Fr* abc = abc_all + (size_t)blockIdx.y * 3 * N;
    Fr zi = *zh_inv;
    Fr a = abc[i], b = abc[(size_t)N + i], c = abc[2 * (size_t)N + i];
    out_all[(size_t)blockIdx.y * N + i] = (a * b - c) * zi;
}

// abc: K x [a | b | c] (each N), clobbered; h: K x N.
template <class Curve>
void qap_h_from_evals(hk_ctx* ctx, uint32_t log_n, uint64_t* d_abc_v, uint64_t* d_h_v, size_t K) {
    using Fr = typename Curve::Fr;
    Fr* abc = reinterpret_cast<Fr*>(d_abc_v);
    Fr* h = reinterpret_cast<Fr*>(d_h_v);
    const size_t N = size_t(1) << log_n;
    HK_REQUIRE(log_n >= 1, HK_ERR_UNSUPPORTED, "qap_h: log_n must be >= 1");
    if (K == 0) return;
    auto dom = ntt_domain<Curve>(ctx, log_n);
    Fr* tmp = ctx->ws_ntt_tmp.as<Fr>((3 + 1) * K * N);
    Fr* ab = tmp + 3 * K * N;      // pointwise result, K x N, contiguous for the batched last transform
    // 3K x ifft (unnormalised, bit-reversed output), in place
    ntt_passes<Curve>(ctx, *dom, false, true, abc, abc, tmp, 3 * K, false, false, nullptr, nullptr, nullptr);
    // 3K x coset fft from bit-reversed coefficients; scale g^rev(p)/N folded into the load
    ntt_passes<Curve>(ctx, *dom, true, false, abc, abc, tmp, 3 * K, false, false,
                      reinterpret_cast<const Fr*>(dom->coset_br), nullptr, nullptr);
    k_qap_pointwise<Fr><<<dim3(ceil_div(N, 256), (unsigned)K), 256, 0, ctx->qstream()>>>(
        abc, (uint32_t)N, reinterpret_cast<const Fr*>(dom->zh_inv), ab);
    HK_LAUNCH_CHECK(ctx);
    // K x coset ifft: DIF(w^-1), last pass scales by g^-j/N (indexed by destination j) and un-permutes
    ntt_passes<Curve>(ctx, *dom, false, true, ab, h, tmp, K, false, true, nullptr,
                      reinterpret_cast<const Fr*>(dom->coset_inv_nat), nullptr, /*src_clobber_ok=*/true);
}

// z: K assignments of (l + nw) elements; abc: K x 3N
template <class Curve>
void r1cs_evals(hk_ctx* ctx, const R1csDev& m, const uint64_t* d_z, uint64_t* d_abc, size_t K) {
    using Fr = typename Curve::Fr;
    const uint32_t N = 1u << m.log_n;
    if (K == 0) return;
    k_r1cs_evals<Fr><<<dim3(ceil_div(N, 128), (unsigned)K), 128, 0, ctx->qstream()>>>(
        m.ptr[0], m.col[0], reinterpret_cast<const Fr*>(m.val[0]), m.ptr[1], m.col[1],
        reinterpret_cast<const Fr*>(m.val[1]), m.ptr[2], m.col[2], reinterpret_cast<const Fr*>(m.val[2]),
        reinterpret_cast<const Fr*>(d_z), (uint32_t)m.n, (uint32_t)m.l, N, reinterpret_cast<Fr*>(d_abc), m.l + m.nw);
    HK_LAUNCH_CHECK(ctx);
}

template void qap_h_from_evals<Bls12381>(hk_ctx*, uint32_t, uint64_t*, uint64_t*, size_t);
template void qap_h_from_evals<Bn254>(hk_ctx*, uint32_t, uint64_t*, uint64_t*, size_t);
template void r1cs_evals<Bls12381>(hk_ctx*, const R1csDev&, const uint64_t*, uint64_t*, size_t);
template void r1cs_evals<Bn254>(hk_ctx*, const R1csDev&, const uint64_t*, uint64_t*, size_t);

std::shared_ptr<R1csDev> r1cs_upload(hk_ctx* ctx, size_t n, size_t l, size_t nw, const uint32_t* const ptr[3],
                                     const uint32_t* const col[3], const uint64_t* const val[3]) {
    auto m = std::make_shared<R1csDev>();
    m->n = n;
    m->l = l;
    m->nw = nw;
    HK_REQUIRE(n + l >= 1, HK_ERR_INVALID, "r1cs: empty");
    uint32_t lg = 0;
    while ((size_t(1) << lg) < n + l) lg++;
    HK_REQUIRE(lg >= 1, HK_ERR_UNSUPPORTED, "r1cs: n + l must be >= 2");
    m->log_n = lg;
    for (int k = 0; k < 3; k++) {
        HK_REQUIRE(ptr[k] && ptr[k][0] == 0, HK_ERR_INVALID, "r1cs: row_ptr must start at 0");
        for (size_t i = 0; i < n; i++)
            HK_REQUIRE(ptr[k][i] <= ptr[k][i + 1], HK_ERR_INVALID, "r1cs: row_ptr must be non-decreasing");
        size_t nnz = ptr[k][n];
        HK_REQUIRE(nnz == 0 || (col[k] && val[k]), HK_ERR_INVALID, "r1cs: col / val missing for a non-empty matrix");
        m->nnz[k] = nnz;
        m->ptr[k] = m->b_ptr[k].as<uint32_t>(n + 1);
        m->col[k] = m->b_col[k].as<uint32_t>(nnz + 1);
        m->val[k] = m->b_val[k].as<uint64_t>(4 * (nnz + 1));
        HK_CUDA(cudaMemcpyAsync(m->ptr[k], ptr[k], (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (nnz) {
            for (size_t j = 0; j < nnz; j++)
                HK_REQUIRE(col[k][j] < l + nw, HK_ERR_INVALID, "r1cs: column index out of range");
            HK_CUDA(cudaMemcpyAsync(m->col[k], col[k], nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
            HK_CUDA(cudaMemcpyAsync(m->val[k], val[k], nnz * 32, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    HK_CUDA(cudaStreamSynchronize(ctx->stream));
    return m;
}

}  // namespace hk
