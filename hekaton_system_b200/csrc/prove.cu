// Fused per-subcircuit CP-Groth16 path on the device.
//   commit : /root/reference/cp-groth16/src/committer.rs:55-98
//   prove  : /root/reference/cp-groth16/src/committer.rs:100-122 -> prover.rs:19-32,53-156
//
// Every O(1) group operation of the reference that has a FIXED base is folded into the
// neighbouring MSM as an extra (base, scalar) pair, so no serial scalar multiplication is
// paid for it:
//   A   = MSM( [a_g[1..], a_g[0], alpha_g, delta_g] , [z[1..], 1, 1, r] )        prover.rs:86-88,158-171
//   B1  = MSM( [b_g[1..], b_g[0], beta_g,  delta_g] , [z[1..], 1, 1, s] )        prover.rs:94-97
//   B2  = MSM( [b_h[1..], b_h[0], beta_h,  delta_h] , [z[1..], 1, 1, s] )  (G2)  prover.rs:106-107
//   L'  = MSM( [ck_last.., delta_g, deltas_g[0..]]  , [w_last.., -rs, -kappa_i] ) prover.rs:117,135,144; committer.rs:113-114
//   H   = MSM( h_g , h[..N-1] )                                                  prover.rs:128-129
//   C   = s*A + r*B1 + L' + H                                                    prover.rs:139-147
// Only s*A and r*B1 (scalar times a freshly computed point) remain as scalar
// multiplications; k_prove_finish runs them on two warps and normalises A, B2, C.
// (r == 0 needs no special case: the reference sets B1 = 0 only to skip work, r*B1 is 0.)
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
#include "prove.cuh"

#include "ntt_impl.cuh"
#include "qap.cuh"

namespace hk {

// ---- small generic kernels --------------------------------------------------------------------
template <class F>
__global__ void k_apply_infinity(Affine<F>* pts, const uint8_t* inf, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (inf[i]) pts[i] = Affine<F>::inf();
}

// `spread` threads per point (only the first of each group works): spread = 32 gives every point its own warp
template <class F>
__global__ void k_jac_to_affine(const Jacobian<F>* in, Affine<F>* out, uint8_t* inf, size_t n, uint32_t spread) {
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g % spread) return;
    size_t i = g / spread;
    if (i >= n) return;
    XYZZ<F> p = in[i].to_xyzz();
    Affine<F> a = p.to_affine();
    out[i] = a;
    if (inf) inf[i] = p.is_inf() ? 1 : 0;
}

// k * P, k a canonical 256-bit integer, 4-bit fixed windows (MSB first)
template <class F>
__device__ __noinline__ XYZZ<F> scalar_mul(const XYZZ<F>& p, const uint64_t k[4]) {
    XYZZ<F> tbl[16];
    tbl[0] = XYZZ<F>::inf();
    tbl[1] = p;
    for (int i = 2; i < 16; i++) {
        tbl[i] = tbl[i - 1];
        tbl[i].add(p);
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = 63; w >= 0; w--) {
        for (int j = 0; j < 4; j++) acc = acc.dbl();
        uint32_t d = (uint32_t)(k[w >> 4] >> ((w & 15) * 4)) & 15u;
        if (d) acc.add(tbl[d]);
    }
    return acc;
}

// the same, four lanes per multiplication (ec.cuh quad_add / quad_dbl): 3.3x shorter dependent chain
template <class F>
__device__ __noinline__ XYZZ<F> quad_scalar_mul(const XYZZ<F>& p, const uint64_t k[4]) {
    XYZZ<F> tbl[16];
    tbl[0] = XYZZ<F>::inf();
    tbl[1] = p;
    for (int i = 2; i < 16; i++) tbl[i] = quad_add(tbl[i - 1], p);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = 63; w >= 0; w--) {
        for (int j = 0; j < 4; j++) acc = quad_dbl(acc);
        uint32_t d = (uint32_t)(k[w >> 4] >> ((w & 15) * 4)) & 15u;
        if (d) acc = quad_add(acc, tbl[d]);
    }
    return acc;
}

// Fixed-base table T[w][d] = d * 16^w * G  (w < 64, d < 16), then out[i] = sum_w T[w][digit_w(k_i)]
template <class F>
__global__ void k_fixed_table(const Affine<F>* gen, XYZZ<F>* tbl) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 64) return;
    XYZZ<F> b = XYZZ<F>::from_affine(*gen);
    for (uint32_t i = 0; i < 4 * w; i++) b = b.dbl();
    XYZZ<F> acc = XYZZ<F>::inf();
    tbl[w * 16] = acc;
    for (int d = 1; d < 16; d++) {
        acc.add(b);
        tbl[w * 16 + d] = acc;
    }
}

template <class F, class Fr>
__global__ void __launch_bounds__(128)
k_fixed_base_mul(const XYZZ<F>* __restrict__ tbl, const Scalar256* __restrict__ sc, int mont, Affine<F>* out,
                 size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k[4] = {sc[i].v[0], sc[i].v[1], sc[i].v[2], sc[i].v[3]};
    if (mont) {
        Fr x;
        for (int j = 0; j < 4; j++) x.r.v[j] = k[j];
        x = x.from_mont();
        for (int j = 0; j < 4; j++) k[j] = x.r.v[j];
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = 0; w < 64; w++) {
        uint32_t d = (uint32_t)(k[w >> 4] >> ((w & 15) * 4)) & 15u;
        if (d) acc.add(tbl[w * 16 + d]);
    }
    out[i] = acc.to_affine();
}

// out[i] = scalars[i] * points[i], normalised -- the reference's `scalar_pairing`
// (/root/reference/distributed-prover/src/pairing_ops.rs:32-39: per-element `*si * *ri`, then `normalize_batch`),
// which the aggregator runs over the A / C / D / input vectors of all proofs (aggregation.rs:236-242,288-310).
// `spread` threads per element (only the first of each group works): 32 = one warp per element, so the
// data-dependent branches of the inversion and the digit tests do not serialise across elements.
template <class F, class Fr>
__global__ void __launch_bounds__(128)
k_scalar_mul_batch(const Affine<F>* __restrict__ pts, const uint8_t* __restrict__ inf_in, const Scalar256* __restrict__ sc,
                   int mont, Affine<F>* out, uint8_t* inf_out, size_t n, uint32_t spread) {
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g % spread) return;
    size_t i = g / spread;
    if (i >= n) return;
    uint64_t k[4] = {sc[i].v[0], sc[i].v[1], sc[i].v[2], sc[i].v[3]};
    if (mont) {
        Fr x;
        for (int j = 0; j < 4; j++) x.r.v[j] = k[j];
        x = x.from_mont();
        for (int j = 0; j < 4; j++) k[j] = x.r.v[j];
    }
    Affine<F> p = pts[i];
    XYZZ<F> acc = (inf_in && inf_in[i]) ? XYZZ<F>::inf() : XYZZ<F>::from_affine(p);
    acc = scalar_mul<F>(acc, k);
    out[i] = acc.to_affine();
    if (inf_out) inf_out[i] = acc.is_inf() ? 1 : 0;
}

// ---- scalar staging ---------------------------------------------------------------------------
// S = from_mont([ z[1..m-1], 1, 1, x ])  (m + 2 canonical scalars)
// blockIdx.y = proof index p in the batch.  Writes both scalar vectors of the proof:
//   out_r[p] = from_mont([z[1..], 1, 1, r_p])   (A)      out_s[p] = from_mont([z[1..], 1, 1, s_p])   (B1, B2)
template <class Fr>
__global__ void k_build_abc_scalars(const Fr* __restrict__ z_all, size_t m, const Fr* r_all, const Fr* s_all,
                                    Scalar256* out_r_all, Scalar256* out_s_all) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m + 2) return;
    const size_t p = blockIdx.y;
    const Fr* z = z_all + p * m;
    Scalar256* out_r = out_r_all + p * (m + 2);
    Scalar256* out_s = out_s_all + p * (m + 2);
    if (i < m + 1) {
        Fr v = (i < m - 1) ? z[i + 1] : Fr::one();
        v = v.from_mont();
        for (int k = 0; k < 4; k++) {
            out_r[i].v[k] = v.r.v[k];
            out_s[i].v[k] = v.r.v[k];
        }
    } else {
        Fr a = r_all[p].from_mont(), b = s_all[p].from_mont();
        for (int k = 0; k < 4; k++) {
            out_r[i].v[k] = a.r.v[k];
            out_s[i].v[k] = b.r.v[k];
        }
    }
}

// dst[p * stride] = from_mont(x[p])
template <class Fr>
__global__ void k_patch_scalar(const Fr* x, Scalar256* dst, size_t stride, size_t K) {
    size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= K) return;
    Fr v = x[p].from_mont();
    for (int k = 0; k < 4; k++) dst[p * stride].v[k] = v.r.v[k];
}

// L' scalars (Montgomery), per proof: [ w_last (nl), -(r s), -kappa_0 .. -kappa_{ns-2} ]
template <class Fr>
__global__ void k_build_l_scalars(const Fr* __restrict__ z_all, size_t m, size_t w_off, size_t nl, const Fr* r,
                                  const Fr* s, const Fr* kappas, uint32_t nk, Fr* out_all) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nl + 1 + nk) return;
    const size_t p = blockIdx.y;
    const Fr* w_last = z_all + p * m + w_off;
    Fr* out = out_all + p * (nl + 1 + nk);
    if (i < nl) out[i] = w_last[i];
    else if (i == nl) out[i] = (r[p] * s[p]).neg();
    else out[i] = kappas[p * nk + (i - nl - 1)].neg();
}

// ---- finish -----------------------------------------------------------------------------------
// Stage 1 (highest-priority lane, right after the A and B1 MSMs; overlapped with the bulk MSMs): the two
// scalar multiplications by a fresh point, and the normalisation of A.  One block per proof p.
//   jacA, jacB1: K each (G1);  r, s: K Montgomery Fr.   sab[2p] = s*A, sab[2p+1] = r*B1
template <class Curve>
__global__ void __launch_bounds__(64)
k_prove_finish_ab(const Jacobian<typename Curve::Fq>* jacA, const Jacobian<typename Curve::Fq>* jacB1,
                  const typename Curve::Fr* r_all, const typename Curve::Fr* s_all, XYZZ<typename Curve::Fq>* sab,
                  Affine<typename Curve::Fq>* outA_all, uint8_t* inf_all) {
    using Fq = typename Curve::Fq;
    const uint32_t p = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane >= 4) return;                      // one quad per scalar multiplication
    if (warp == 0) {
        XYZZ<Fq> a = jacA[p].to_xyzz();
        auto s = s_all[p].from_mont();
        XYZZ<Fq> sa = quad_scalar_mul<Fq>(a, s.r.v);
        if (lane == 0) {
            sab[2 * p] = sa;
            outA_all[p] = a.to_affine();
            inf_all[3 * p + 0] = a.is_inf() ? 1 : 0;
        }
    } else {
        auto r = r_all[p].from_mont();
        XYZZ<Fq> rb = quad_scalar_mul<Fq>(jacB1[p].to_xyzz(), r.r.v);
        if (lane == 0) sab[2 * p + 1] = rb;
    }
}

// Normalisation of B2 (G2), on the lane of the G2 MSM.  One warp per proof (lane 0 works) so the
// K inversions do not serialise each other's divergent branches.
template <class Curve>
__global__ void __launch_bounds__(32)
k_prove_finish_b2(const Jacobian<typename Curve::Fq2>* jacB2, Affine<typename Curve::Fq2>* outB_all, uint8_t* inf_all) {
    using Fq2 = typename Curve::Fq2;
    const uint32_t p = blockIdx.x;
    if (threadIdx.x != 0) return;
    XYZZ<Fq2> b = jacB2[p].to_xyzz();
    outB_all[p] = b.to_affine();
    inf_all[3 * p + 1] = b.is_inf() ? 1 : 0;
}

// Stage 2 (main stream, after every lane): C = s*A + r*B1 + L' + H -> affine.  One warp per proof
// (lane 0 works: the inversion's data-dependent branches must not serialise across proofs).
template <class Curve>
__global__ void __launch_bounds__(32)
k_prove_finish_c(const XYZZ<typename Curve::Fq>* sab, const Jacobian<typename Curve::Fq>* jacL,
                 const Jacobian<typename Curve::Fq>* jacH, Affine<typename Curve::Fq>* outC_all, uint8_t* inf_all,
                 uint32_t K) {
    using Fq = typename Curve::Fq;
    const uint32_t p = blockIdx.x;
    if (p >= K || threadIdx.x != 0) return;
    XYZZ<Fq> c = sab[2 * p];
    c.add(sab[2 * p + 1]);
    c.add(jacL[p].to_xyzz());
    c.add(jacH[p].to_xyzz());
    outC_all[p] = c.to_affine();
    inf_all[3 * p + 2] = c.is_inf() ? 1 : 0;
}

// ---- host orchestration -----------------------------------------------------------------------
template <class Curve>
static void bases_finish_upload(hk_ctx* ctx, BaseSet& bs, const uint8_t* d_inf) {
    if (!d_inf || bs.n == 0) return;
    if (bs.group == 1)
        k_apply_infinity<typename Curve::Fq><<<ceil_div(bs.n, 256), 256, 0, ctx->stream>>>(
            reinterpret_cast<Affine<typename Curve::Fq>*>(bs.d), d_inf, bs.n);
    else
        k_apply_infinity<typename Curve::Fq2><<<ceil_div(bs.n, 256), 256, 0, ctx->stream>>>(
            reinterpret_cast<Affine<typename Curve::Fq2>*>(bs.d), d_inf, bs.n);
    HK_LAUNCH_CHECK(ctx);
}

template <class Curve>
void apply_infinity(hk_ctx* ctx, BaseSet& bs, const uint8_t* d_inf) {
    bases_finish_upload<Curve>(ctx, bs, d_inf);
}
template void apply_infinity<Bls12381>(hk_ctx*, BaseSet&, const uint8_t*);
template void apply_infinity<Bn254>(hk_ctx*, BaseSet&, const uint8_t*);

template <class Curve, int GROUP>
void jac_to_affine(hk_ctx* ctx, const void* d_jac, void* d_aff, uint8_t* d_inf, size_t n, cudaStream_t st) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    if (n == 0) return;
    if (!st) st = ctx->stream;
    // few points: one warp each (the inversion's data-dependent branches do not serialise across points)
    if (n <= 4096)
        k_jac_to_affine<F><<<(unsigned)n, 32, 0, st>>>(reinterpret_cast<const Jacobian<F>*>(d_jac),
                                                       reinterpret_cast<Affine<F>*>(d_aff), d_inf, n, 32);
    else
        k_jac_to_affine<F><<<ceil_div(n, 64), 64, 0, st>>>(reinterpret_cast<const Jacobian<F>*>(d_jac),
                                                           reinterpret_cast<Affine<F>*>(d_aff), d_inf, n, 1);
    HK_LAUNCH_CHECK(ctx);
}
template void jac_to_affine<Bls12381, 1>(hk_ctx*, const void*, void*, uint8_t*, size_t, cudaStream_t);
template void jac_to_affine<Bls12381, 2>(hk_ctx*, const void*, void*, uint8_t*, size_t, cudaStream_t);
template void jac_to_affine<Bn254, 1>(hk_ctx*, const void*, void*, uint8_t*, size_t, cudaStream_t);
template void jac_to_affine<Bn254, 2>(hk_ctx*, const void*, void*, uint8_t*, size_t, cudaStream_t);

template <class Curve, int GROUP>
void fixed_base_mul(hk_ctx* ctx, const void* d_gen_affine, const Scalar256* d_scalars, size_t n, bool mont,
                    void* d_out_affine) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    using Fr = typename Curve::Fr;
    if (n == 0) return;
    XYZZ<F>* tbl = ctx->ws_misc.as<XYZZ<F>>(64 * 16);
    k_fixed_table<F><<<2, 32, 0, ctx->stream>>>(reinterpret_cast<const Affine<F>*>(d_gen_affine), tbl);
    HK_LAUNCH_CHECK(ctx);
    k_fixed_base_mul<F, Fr><<<ceil_div(n, 128), 128, 0, ctx->stream>>>(tbl, d_scalars, mont ? 1 : 0,
                                                                       reinterpret_cast<Affine<F>*>(d_out_affine), n);
    HK_LAUNCH_CHECK(ctx);
}
template void fixed_base_mul<Bls12381, 1>(hk_ctx*, const void*, const Scalar256*, size_t, bool, void*);
template void fixed_base_mul<Bls12381, 2>(hk_ctx*, const void*, const Scalar256*, size_t, bool, void*);
template void fixed_base_mul<Bn254, 1>(hk_ctx*, const void*, const Scalar256*, size_t, bool, void*);
template void fixed_base_mul<Bn254, 2>(hk_ctx*, const void*, const Scalar256*, size_t, bool, void*);

template <class Curve, int GROUP>
void scalar_mul_batch(hk_ctx* ctx, const void* d_pts, const uint8_t* d_inf, const Scalar256* d_scalars, size_t n, bool mont,
                      void* d_out_affine, uint8_t* d_inf_out) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    using Fr = typename Curve::Fr;
    if (n == 0) return;
    const uint32_t spread = n <= 65536 ? 32u : 1u;
    k_scalar_mul_batch<F, Fr><<<ceil_div(n * spread, 128), 128, 0, ctx->stream>>>(
        reinterpret_cast<const Affine<F>*>(d_pts), d_inf, d_scalars, mont ? 1 : 0, reinterpret_cast<Affine<F>*>(d_out_affine),
        d_inf_out, n, spread);
    HK_LAUNCH_CHECK(ctx);
}
template void scalar_mul_batch<Bls12381, 1>(hk_ctx*, const void*, const uint8_t*, const Scalar256*, size_t, bool, void*, uint8_t*);
template void scalar_mul_batch<Bls12381, 2>(hk_ctx*, const void*, const uint8_t*, const Scalar256*, size_t, bool, void*, uint8_t*);
template void scalar_mul_batch<Bn254, 1>(hk_ctx*, const void*, const uint8_t*, const Scalar256*, size_t, bool, void*, uint8_t*);
template void scalar_mul_batch<Bn254, 2>(hk_ctx*, const void*, const uint8_t*, const Scalar256*, size_t, bool, void*, uint8_t*);

static std::shared_ptr<BaseSet> get_bases_sp(hk_ctx* ctx, uint64_t h, int group) {
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    auto it = ctx->sh->bases.find(h);
    HK_REQUIRE(it != ctx->sh->bases.end(), HK_ERR_HANDLE, "unknown base-set handle");
    HK_REQUIRE(it->second->group == group, HK_ERR_INVALID, "base set is in the wrong group");
    return it->second;
}
static BaseSet& get_bases(hk_ctx* ctx, uint64_t h, int group) { return *get_bases_sp(ctx, h, group); }

static void alloc_set(BaseSet& bs, int group, size_t n, size_t point_bytes) {
    bs.group = group;
    bs.n = n;
    bs.point_bytes = point_bytes;
    cudaError_t e = cudaMalloc(&bs.d, n * point_bytes + 256);
    if (e != cudaSuccess) throw HkError{HK_ERR_NOMEM, std::string("cudaMalloc(pk bases): ") + cudaGetErrorString(e)};
}

template <class Curve>
std::shared_ptr<PkDev> pk_create(hk_ctx* ctx, const hk_pk_desc& d) {
    using Fq = typename Curve::Fq;
    using Fq2 = typename Curve::Fq2;
    const size_t P1 = sizeof(Affine<Fq>), P2 = sizeof(Affine<Fq2>);
    auto pk = std::make_shared<PkDev>();
    HK_REQUIRE(d.num_stages >= 1 && d.num_stages <= 8, HK_ERR_INVALID, "pk: num_stages must be 1..8");
    {
        std::lock_guard<std::mutex> lk(ctx->sh->mu);
        auto rit = ctx->sh->r1cs.find(d.r1cs);
        HK_REQUIRE(rit != ctx->sh->r1cs.end(), HK_ERR_HANDLE, "pk: unknown r1cs handle");
        pk->r1cs = rit->second;
    }
    pk->num_stages = d.num_stages;
    const size_t l = pk->r1cs->l, nw = pk->r1cs->nw, m = l + nw;
    pk->m = m;
    const size_t N = size_t(1) << pk->r1cs->log_n;
    BaseSet& a_g = get_bases(ctx, d.a_g, 1);
    BaseSet& b_g = get_bases(ctx, d.b_g, 1);
    BaseSet& b_h = get_bases(ctx, d.b_h, 2);
    std::shared_ptr<BaseSet> h_g_sp = get_bases_sp(ctx, d.h_g, 1);
    BaseSet& h_g = *h_g_sp;
    BaseSet& deltas = get_bases(ctx, d.deltas_g, 1);
    HK_REQUIRE(a_g.n == m && b_g.n == m && b_h.n == m, HK_ERR_LENGTH, "pk: a_g/b_g/b_h must have one point per variable");
    HK_REQUIRE(h_g.n == N - 1, HK_ERR_LENGTH, "pk: h_g must have N-1 points (prover.rs:128)");
    HK_REQUIRE(deltas.n == d.num_stages, HK_ERR_LENGTH, "pk: deltas_g must have num_stages points");
    HK_REQUIRE(m >= 1, HK_ERR_INVALID, "pk: no variables");
    for (uint32_t i = 0; i <= d.num_stages; i++) pk->stage_start.push_back(d.stage_start[i]);
    HK_REQUIRE(pk->stage_start[0] == 0 && pk->stage_start[d.num_stages] <= nw, HK_ERR_INVALID,
               "pk: stage ranges must start at 0 and stay within the witness");
    cudaStream_t st = ctx->stream;
    auto cp = [&](void* dst, const void* src, size_t bytes, cudaMemcpyKind k) {
        if (bytes) HK_CUDA(cudaMemcpyAsync(dst, src, bytes, k, st));
    };
    // [q[1..m-1], q[0], vk_param, delta]
    alloc_set(pk->a_ext, 1, m + 2, P1);
    alloc_set(pk->b1_ext, 1, m + 2, P1);
    alloc_set(pk->b2_ext, 2, m + 2, P2);
    char* a_ext = (char*)pk->a_ext.d;
    cp(a_ext, (char*)a_g.d + P1, (m - 1) * P1, cudaMemcpyDeviceToDevice);
    cp(a_ext + (m - 1) * P1, a_g.d, P1, cudaMemcpyDeviceToDevice);
    cp(a_ext + m * P1, d.alpha_g, P1, cudaMemcpyHostToDevice);
    cp(a_ext + (m + 1) * P1, (char*)deltas.d + (d.num_stages - 1) * P1, P1, cudaMemcpyDeviceToDevice);
    char* b1_ext = (char*)pk->b1_ext.d;
    cp(b1_ext, (char*)b_g.d + P1, (m - 1) * P1, cudaMemcpyDeviceToDevice);
    cp(b1_ext + (m - 1) * P1, b_g.d, P1, cudaMemcpyDeviceToDevice);
    cp(b1_ext + m * P1, d.beta_g, P1, cudaMemcpyHostToDevice);
    cp(b1_ext + (m + 1) * P1, (char*)deltas.d + (d.num_stages - 1) * P1, P1, cudaMemcpyDeviceToDevice);
    char* b2_ext = (char*)pk->b2_ext.d;
    cp(b2_ext, (char*)b_h.d + P2, (m - 1) * P2, cudaMemcpyDeviceToDevice);
    cp(b2_ext + (m - 1) * P2, b_h.d, P2, cudaMemcpyDeviceToDevice);
    cp(b2_ext + m * P2, d.beta_h, P2, cudaMemcpyHostToDevice);
    cp(b2_ext + (m + 1) * P2, d.last_delta_h, P2, cudaMemcpyHostToDevice);
    // per-stage committer keys: [ck[stage].., delta_last]; the last stage's doubles as L' with the
    // deltas appended: [ck_last.., delta_last, deltas_g[0..ns-2]]
    for (uint32_t sidx = 0; sidx < d.num_stages; sidx++) {
        BaseSet& ck = get_bases(ctx, d.ck[sidx], 1);
        size_t want = pk->stage_start[sidx + 1] - pk->stage_start[sidx];
        HK_REQUIRE(ck.n == want, HK_ERR_LENGTH, "pk: ck[stage] length must equal the stage's witness count");
        pk->ck_ext.emplace_back(new BaseSet());
        size_t extra = (sidx + 1 == d.num_stages) ? d.num_stages : 1;
        alloc_set(*pk->ck_ext.back(), 1, ck.n + extra, P1);
        char* e = (char*)pk->ck_ext.back()->d;
        cp(e, ck.d, ck.n * P1, cudaMemcpyDeviceToDevice);
        cp(e + ck.n * P1, (char*)deltas.d + (d.num_stages - 1) * P1, P1, cudaMemcpyDeviceToDevice);
        if (extra > 1) cp(e + (ck.n + 1) * P1, deltas.d, (d.num_stages - 1) * P1, cudaMemcpyDeviceToDevice);
    }
    pk->h_g = h_g_sp;
    pk->h_g_handle = d.h_g;
    HK_CUDA(cudaStreamSynchronize(st));
    // window-shifted tables for every base vector of the key (one-off; W x the memory)
    // The assignment-scalar MSMs (A, B1, B2, L, commitments) see mostly-bit witnesses -> smaller
    // windows; the H MSM sees uniform scalars.
    const double wd = d.witness_density_permille ? d.witness_density_permille / 1000.0 : 0.15;
    bases_precompute<Curve, 1>(ctx, pk->a_ext, wd);
    bases_precompute<Curve, 1>(ctx, pk->b1_ext, wd);
    bases_precompute<Curve, 2>(ctx, pk->b2_ext, wd);
    for (auto& ck : pk->ck_ext) bases_precompute<Curve, 1>(ctx, *ck, wd);
    bases_precompute<Curve, 1>(ctx, *pk->h_g, 1.0);
    // B in G1 and B in G2 take the same scalars: when the two base sets have the same table geometry and the same
    // points at infinity (b_i(tau) = 0 in both groups -- true for every honestly generated key), one sorted digit
    // list serves both accumulations
    pk->share_b_sort = pk->b1_ext.tbl && pk->b2_ext.tbl && pk->b1_ext.tbl_c == pk->b2_ext.tbl_c &&
                       pk->b1_ext.tbl_W == pk->b2_ext.tbl_W && pk->b1_ext.n == pk->b2_ext.n;
    if (pk->share_b_sort) {
        std::vector<uint8_t> s1(pk->b1_ext.n), s2(pk->b2_ext.n);
        HK_CUDA(cudaMemcpy(s1.data(), pk->b1_ext.skip, s1.size(), cudaMemcpyDeviceToHost));
        HK_CUDA(cudaMemcpy(s2.data(), pk->b2_ext.skip, s2.size(), cudaMemcpyDeviceToHost));
        pk->share_b_sort = s1 == s2;
    }
    return pk;
}
template std::shared_ptr<PkDev> pk_create<Bls12381>(hk_ctx*, const hk_pk_desc&);
template std::shared_ptr<PkDev> pk_create<Bn254>(hk_ctx*, const hk_pk_desc&);

// commit scalars (Montgomery), per proof p (blockIdx.y): [ w_stage (nw), kappa_p ]
template <class Fr>
__global__ void k_build_commit_scalars(const Fr* __restrict__ z_all, size_t z_stride, size_t w_off, size_t nw,
                                       const Fr* kappa, Fr* out_all) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i > nw) return;
    const size_t p = blockIdx.y;
    const Fr* w = z_all + p * z_stride + w_off;
    out_all[p * (nw + 1) + i] = (i < nw) ? w[i] : kappa[p];
}

// K commitments of one stage in one pipeline.  z: K assignments, `z_stride` elements apart, each at
// least l + stage_start[stage+1] long; kappa: K; out: K affine points + K infinity flags.
template <class Curve>
void commit_run(hk_ctx* ctx, PkDev& pk, uint32_t stage, size_t K, const uint64_t* d_z, size_t z_stride,
                const uint64_t* d_kappa, void* d_out_affine, uint8_t* d_inf) {
    using Fr = typename Curve::Fr;
    using Fq = typename Curve::Fq;
    HK_REQUIRE(stage < pk.num_stages, HK_ERR_INVALID, "commit: no such stage (\"no more values left in committing key\")");
    if (K == 0) return;
    NvtxRange nvtx_commit("Groth16::Commit");                  // committer.rs:59
    const size_t l = pk.r1cs->l;
    const size_t nw = pk.stage_start[stage + 1] - pk.stage_start[stage];
    // a commitment is a small latency-bound MSM: run it on a prioritised lane so that, when other
    // contexts keep the GPU busy (proofs of the previous batch), its kernels are served promptly
    cudaStream_t ls = ctx->lane_stream(1);
    HK_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    HK_CUDA(cudaStreamWaitEvent(ls, ctx->ev_fork, 0));
    Fr* sc = ctx->ws_io[0].as<Fr>(K * (nw + 1));
    k_build_commit_scalars<Fr><<<dim3(ceil_div(nw + 1, 256), (unsigned)K), 256, 0, ls>>>(
        reinterpret_cast<const Fr*>(d_z), z_stride, l + pk.stage_start[stage], nw, reinterpret_cast<const Fr*>(d_kappa), sc);
    HK_LAUNCH_CHECK(ctx);
    Jacobian<Fq>* jac = ctx->ws_out.as<Jacobian<Fq>>(8 * K);
    msm_run<Curve, 1>(ctx, msm_view(*pk.ck_ext[stage]), 0, nw + 1, reinterpret_cast<const Scalar256*>(sc), K, true, jac, 0, 1);
    jac_to_affine<Curve, 1>(ctx, jac, d_out_affine, d_inf, K, ls);
    HK_CUDA(cudaEventRecord(ctx->ev_lane_done[1], ls));
    HK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane_done[1], 0));
}
template void commit_run<Bls12381>(hk_ctx*, PkDev&, uint32_t, size_t, const uint64_t*, size_t, const uint64_t*, void*,
                                   uint8_t*);
template void commit_run<Bn254>(hk_ctx*, PkDev&, uint32_t, size_t, const uint64_t*, size_t, const uint64_t*, void*, uint8_t*);

template <class Curve>
void prove_run(hk_ctx* ctx, PkDev& pk, size_t K, const uint64_t* d_z_v, const uint64_t* d_r, const uint64_t* d_s,
               const uint64_t* d_kappas, void* d_outA, void* d_outB, void* d_outC, uint8_t* d_inf) {
    using Fr = typename Curve::Fr;
    using Fq = typename Curve::Fq;
    using Fq2 = typename Curve::Fq2;
    if (K == 0) return;
    NvtxRange nvtx_prover("Groth16::Prover");                  // prover.rs:58
    const Fr* z = reinterpret_cast<const Fr*>(d_z_v);
    const Fr* r = reinterpret_cast<const Fr*>(d_r);
    const Fr* s = reinterpret_cast<const Fr*>(d_s);
    const size_t m = pk.m, l = pk.r1cs->l;
    const uint32_t log_n = pk.r1cs->log_n;
    const size_t N = size_t(1) << log_n;
    const uint32_t ns = pk.num_stages;
    cudaStream_t st = ctx->stream;
    const unsigned Ku = (unsigned)K;
    Jacobian<Fq>* jac = ctx->ws_out.as<Jacobian<Fq>>(8 * K);        // [A | B1 | L' | H] x K, then B2 x K
    Jacobian<Fq2>* jacB2 = reinterpret_cast<Jacobian<Fq2>*>(jac + 4 * K);
    // A uses [z.., 1, 1, r]; B1 and B2 use [z.., 1, 1, s]
    Scalar256* Sr = ctx->ws_io[0].as<Scalar256>(K * (m + 2));
    Scalar256* Ss = ctx->ws_io[4].as<Scalar256>(K * (m + 2));
    k_build_abc_scalars<Fr><<<dim3(ceil_div(m + 2, 256), Ku), 256, 0, st>>>(z, m, r, s, Sr, Ss);
    HK_LAUNCH_CHECK(ctx);
    // Lanes (common.cuh): every MSM gets its own lane, so all sorts start at once, the accumulation
    // kernels are served back to back in lane order and each MSM's latency-bound tail (folds, bucket
    // reduction) hides under the next accumulation.  Order: A and B1 first, because the 5 ms of
    // two-warp scalar multiplications s*A, r*B1 depend on them and must be hidden under the bulk; then
    // the G2 MSM (longest tail), the H MSM (half of all additions; its scalars need the witness map,
    // which runs on the lane's control stream meanwhile), and L' last (shortest tail left exposed).
    enum { LA = 1, LB1 = 2, LB2 = 3, LH = 4, LL = 5 };
    HK_CUDA(cudaEventRecord(ctx->ev_fork, st));
    for (int ln = 1; ln < hk_ctx::LANES; ln++) HK_CUDA(cudaStreamWaitEvent(ctx->lane_stream(ln), ctx->ev_fork, 0));
    {
        NvtxRange r_("Compute A");                             // prover.rs:84
        msm_run<Curve, 1>(ctx, msm_view(pk.a_ext), 0, m + 2, Sr, K, false, jac + 0, 0, LA);
    }
    HK_CUDA(cudaEventRecord(ctx->ev_lane_done[LA], ctx->lane_stream(LA)));
    MsmSorted sorted_s;
    nvtxRangePushA("Compute B in G1");                         // prover.rs:94
    if (pk.share_b_sort) {
        sorted_s = msm_sort<Curve>(ctx, msm_view(pk.b1_ext), 0, m + 2, Ss, K, false, 0, LB1);
        msm_accumulate<Curve, 1>(ctx, sorted_s, msm_view(pk.b1_ext), jac + K, LB1);
    } else {
        msm_run<Curve, 1>(ctx, msm_view(pk.b1_ext), 0, m + 2, Ss, K, false, jac + K, 0, LB1);
    }
    nvtxRangePop();
    // s*A, r*B1 and A's normalisation, on B1's control stream once A is there too
    XYZZ<Fq>* sab = ctx->ws_io[2].as<XYZZ<Fq>>(2 * K);
    HK_CUDA(cudaStreamWaitEvent(ctx->lane_stream(LB1), ctx->ev_lane_done[LA], 0));
    k_prove_finish_ab<Curve><<<Ku, 64, 0, ctx->lane_stream(LB1)>>>(jac, jac + K, r, s, sab,
                                                                 reinterpret_cast<Affine<Fq>*>(d_outA), d_inf);
    HK_LAUNCH_CHECK(ctx);
    HK_CUDA(cudaEventRecord(ctx->ev_lane_done[LB1], ctx->lane_stream(LB1)));
    // the G2 MSM + B2's normalisation
    nvtxRangePushA("Compute B in G2");                         // prover.rs:105
    if (pk.share_b_sort)
        msm_accumulate<Curve, 2>(ctx, sorted_s, msm_view(pk.b2_ext), jacB2, LB2);     // same scalars, same list
    else
        msm_run<Curve, 2>(ctx, msm_view(pk.b2_ext), 0, m + 2, Ss, K, false, jacB2, 0, LB2);
    k_prove_finish_b2<Curve><<<Ku, 32, 0, ctx->lane_stream(LB2)>>>(jacB2, reinterpret_cast<Affine<Fq2>*>(d_outB), d_inf);
    HK_LAUNCH_CHECK(ctx);
    HK_CUDA(cudaEventRecord(ctx->ev_lane_done[LB2], ctx->lane_stream(LB2)));
    nvtxRangePop();
    // witness map, then the H MSM (they need only z)
    uint64_t* abc = ctx->ws_qap[0].as<uint64_t>(4 * 3 * N * K);
    uint64_t* h = ctx->ws_qap[1].as<uint64_t>(4 * N * K);
    ntt_domain<Curve>(ctx, log_n);                 // build (and synchronise) the tables before switching streams
    ctx->qap_stream = ctx->lane_stream(LH);
    nvtxRangePushA("R1CS to QAP witness map");                 // prover.rs:122
    try {
        r1cs_evals<Curve>(ctx, *pk.r1cs, d_z_v, abc, K);
        qap_h_from_evals<Curve>(ctx, log_n, abc, h, K);
    } catch (...) {
        ctx->qap_stream = nullptr;
        nvtxRangePop();
        throw;
    }
    ctx->qap_stream = nullptr;
    nvtxRangePop();
    {
        NvtxRange r_("Compute H");                             // prover.rs:127
        msm_run<Curve, 1>(ctx, msm_view(*pk.h_g), 0, N - 1, reinterpret_cast<const Scalar256*>(h), K, true, jac + 3 * K,
                          /*sc_stride=*/N, LH);
    }
    HK_CUDA(cudaEventRecord(ctx->ev_lane_done[LH], ctx->lane_stream(LH)));
    // L'
    NvtxRange nvtx_l("Compute L");                             // prover.rs:112
    const size_t nl = pk.stage_start[ns] - pk.stage_start[ns - 1];
    Fr* Ls = ctx->ws_io[1].as<Fr>(K * (nl + ns));
    k_build_l_scalars<Fr><<<dim3(ceil_div(nl + ns, 256), Ku), 256, 0, ctx->lane_stream(LL)>>>(
        z, m, l + pk.stage_start[ns - 1], nl, r, s, reinterpret_cast<const Fr*>(d_kappas), ns - 1, Ls);
    HK_LAUNCH_CHECK(ctx);
    msm_run<Curve, 1>(ctx, msm_view(*pk.ck_ext[ns - 1]), 0, nl + ns, reinterpret_cast<const Scalar256*>(Ls), K, true,
                      jac + 2 * K, 0, LL);
    // C = s*A + r*B1 + L' + H on L''s control stream (the caller's stream usually has the lowest priority
    // and would be served only after every queued accumulation of the batches behind this one)
    NvtxRange nvtx_c("Finish C");                              // prover.rs:140
    cudaStream_t fin = ctx->lane_stream(LL);
    for (int ln : {LB1, LH}) HK_CUDA(cudaStreamWaitEvent(fin, ctx->ev_lane_done[ln], 0));
    k_prove_finish_c<Curve><<<Ku, 32, 0, fin>>>(sab, jac + 2 * K, jac + 3 * K, reinterpret_cast<Affine<Fq>*>(d_outC), d_inf,
                                                Ku);
    HK_LAUNCH_CHECK(ctx);
    HK_CUDA(cudaEventRecord(ctx->ev_lane_done[LL], fin));
    for (int ln : {LB2, LL}) HK_CUDA(cudaStreamWaitEvent(st, ctx->ev_lane_done[ln], 0));
}
template void prove_run<Bls12381>(hk_ctx*, PkDev&, size_t, const uint64_t*, const uint64_t*, const uint64_t*,
                                  const uint64_t*, void*, void*, void*, uint8_t*);
template void prove_run<Bn254>(hk_ctx*, PkDev&, size_t, const uint64_t*, const uint64_t*, const uint64_t*,
                               const uint64_t*, void*, void*, void*, uint8_t*);

}  // namespace hk
