// Radix-2 NTT over Fr for sm_100a.  Replaces ark-poly 0.4 `Radix2EvaluationDomain::
// {fft_in_place, ifft_in_place}` (+ `get_coset(F::GENERATOR)`), which the reference reaches through
// `QAP::witness_map` at /root/reference/cp-groth16/src/prover.rs:123.
//
// Domain conventions: size N = 2^log_n, group_gen = TWO_ADIC_ROOT^(2^(S-log_n)),
// forward out[k] = sum_j x_j w^(jk) (natural order in and out), inverse includes 1/N, coset
// forward pre-multiplies x_i by g^i, coset inverse post-multiplies by g^-i.
//
// Kernel structure: a transform is 1-3 passes; each pass runs up to 11 butterfly stages on a
// 2^T-element tile staged in shared memory (two 16-byte half-element planes so that unit-stride
// LDS.128/STS.128 are conflict free), reading and writing global memory once.  Tiles of the
// strided passes take 2^q consecutive elements per row (q = 2: 128-byte contiguous runs).
//   DIF pass: stages ascending, (u,v) -> (u+v, (u-v)w)   natural in -> bit-reversed out
//   DIT pass: stages descending, (u,v) -> (u+vw, u-vw)   bit-reversed in -> natural out
// Bit-reversal, coset scaling and 1/N are folded into the first/last pass (PERM_IN/PERM_OUT,
// pre/post scale tables).  Twiddles come from a per-stage compacted table
// tw[off_s + j] = root^(j 2^s), off_s = N - (N >> s), so stage s reads a contiguous range.
#pragma once
#include "common.cuh"

namespace hk {

template <class Curve>
struct FrConsts;
template <>
struct FrConsts<Bls12381> {
    using Fr = Bls12381::Fr;
    __device__ static Fr root() {
        Fr x;
        for (int i = 0; i < 4; i++) x.r.v[i] = bls12_381_fr_ROOT[i];
        return x;
    }
    __device__ static Fr gen() {
        Fr x;
        for (int i = 0; i < 4; i++) x.r.v[i] = bls12_381_fr_GEN[i];
        return x;
    }
};
template <>
struct FrConsts<Bn254> {
    using Fr = Bn254::Fr;
    __device__ static Fr root() {
        Fr x;
        for (int i = 0; i < 4; i++) x.r.v[i] = bn254_fr_ROOT[i];
        return x;
    }
    __device__ static Fr gen() {
        Fr x;
        for (int i = 0; i < 4; i++) x.r.v[i] = bn254_fr_GEN[i];
        return x;
    }
};

static constexpr uint32_t NTT_TMAX = 11;   // tile = 2^11 elements = 64 KiB of shared memory
static constexpr uint32_t NTT_QMAX = 2;
static constexpr uint32_t NTT_TPB = 256;   // one radix-8 group per thread and round on a full tile

enum : uint32_t { NTT_PERM_IN = 1, NTT_PERM_OUT = 2, NTT_TOP_PASS = 4, NTT_RMAX_SHIFT = 4 };   // bits 4-5: stages per round (0 = 3)

struct NttDomain {
    uint32_t log_n = 0;
    DevBuf buf;
    // all tables hold N (or N-1) Fr elements
    void* tw_fwd = nullptr;        // root = w
    void* tw_inv = nullptr;        // root = w^-1
    void* coset_nat = nullptr;     // g^i
    void* coset_inv_nat = nullptr; // g^-i / N
    void* coset_br = nullptr;      // g^rev(p) / N          (pipeline: unnormalised DIF output -> coset DIT input)
    void* coset_inv_br = nullptr;  // g^-rev(p) / N         (pipeline: last DIF output -> natural h)
    void* n_inv = nullptr;         // single element 1/N
    void* zh_inv = nullptr;        // single element 1/(g^N - 1)
};

struct NttPass {
    uint32_t s0, k, q, T;
};

// `perm`: the transform carries a bit reversal on its L = 0 pass (hk_ntt: natural order in and out).  A tile of
// 2^11 CONSECUTIVE elements then gathers / scatters single 32-byte elements 2^(log_n - 11) apart, which
// over-fetches DRAM 2-4x once the vector outgrows L2.  So for log_n >= NTT_TOP_MIN that pass takes only 9
// stages and its tile carries the TOP two index bits as passengers: after the bit reversal they are the two
// lowest address bits, i.e. every access is a 128-byte run.  Same number of passes for log_n >= 21.
static constexpr uint32_t NTT_TOP_MIN = 21;
static constexpr uint32_t NTT_TOP_BITS = 2;

inline std::vector<NttPass> ntt_plan(uint32_t log_n, bool perm = false, uint32_t tmax = NTT_TMAX) {
    std::vector<NttPass> p;
    if (log_n == 0) return p;
    if (log_n <= tmax) {
        p.push_back({0, log_n, 0, log_n});
        return p;
    }
    const bool top = perm && log_n >= NTT_TOP_MIN;
    uint32_t last = top ? tmax - NTT_TOP_BITS : tmax;
    uint32_t rest = log_n - last;
    uint32_t per = tmax - NTT_QMAX;
    uint32_t np = (rest + per - 1) / per;
    uint32_t s = 0;
    for (uint32_t i = 0; i < np; i++) {
        uint32_t k = rest / np + (i < rest % np ? 1 : 0);
        uint32_t L = log_n - s - k;
        uint32_t q = L < NTT_QMAX ? L : NTT_QMAX;
        p.push_back({s, k, q, tmax});
        s += k;
    }
    p.push_back({s, last, 0, tmax});
    return p;
}

// Tile size for K transforms of 2^log_n elements.  A pass is one CTA per tile, and a CTA walks its tile's stages in
// sequence (~40 us for 11 stages on 2^11 elements), so a grid that cannot fill the 148 SMs is pure latency: a single
// 2^16 transform is 32 tiles.  When 2^11-element tiles give fewer than two CTAs per SM, take the smallest tile
// (down to 2^9) that needs no more passes: 2^16 -> 2^9 (128 tiles), 2^17 / 2^18 -> 2^10.  Such a pass also runs
// fewer stages per register round (1 on 2^9 tiles, 2 on 2^10) with 256 threads: a radix-8 round is 12 multiplications
// per thread in three dependent steps, a radix-2 round one -- the chain a lone warp walks is what such a launch costs.
inline uint32_t ntt_tile_bits(uint32_t log_n, size_t K, int sm_count) {
    if (log_n <= 9 || log_n >= NTT_TOP_MIN) return NTT_TMAX;
    const size_t tiles = log_n > NTT_TMAX ? (K << (log_n - NTT_TMAX)) : K;
    if (tiles >= (size_t)2 * sm_count) return NTT_TMAX;
    auto passes = [log_n](uint32_t t) -> uint32_t {
        return log_n <= t ? 1u : 1u + (log_n - t + (t - NTT_QMAX) - 1) / (t - NTT_QMAX);
    };
    uint32_t t = NTT_TMAX;
    while (t > 9 && passes(t - 1) == passes(NTT_TMAX)) t--;
    return t;
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

// shared-memory element access: two planes of 16-byte halves.  The element index is swizzled
// (e ^ ((e >> 3) & 7)) so that both the unit-stride accesses of the load/store phases and the
// power-of-two-strided accesses of the register-radix rounds spread over all banks: the 8 threads of a
// quarter-warp always touch 8 different 16-byte bank groups.
__device__ __forceinline__ uint32_t sm_swz(uint32_t e) { return e ^ ((e >> 3) & 7u); }
template <class Fr>
__device__ __forceinline__ Fr sm_load(const uint4* lo, const uint4* hi, uint32_t e) {
    e = sm_swz(e);
    uint4 a = lo[e], b = hi[e];
    Fr x;
    x.r.v[0] = (uint64_t)a.x | ((uint64_t)a.y << 32);
    x.r.v[1] = (uint64_t)a.z | ((uint64_t)a.w << 32);
    x.r.v[2] = (uint64_t)b.x | ((uint64_t)b.y << 32);
    x.r.v[3] = (uint64_t)b.z | ((uint64_t)b.w << 32);
    return x;
}
template <class Fr>
__device__ __forceinline__ void sm_store(uint4* lo, uint4* hi, uint32_t e, const Fr& x) {
    e = sm_swz(e);
    lo[e] = make_uint4((uint32_t)x.r.v[0], (uint32_t)(x.r.v[0] >> 32), (uint32_t)x.r.v[1], (uint32_t)(x.r.v[1] >> 32));
    hi[e] = make_uint4((uint32_t)x.r.v[2], (uint32_t)(x.r.v[2] >> 32), (uint32_t)x.r.v[3], (uint32_t)(x.r.v[3] >> 32));
}

// tile element e -> global index (see k_ntt_pass)
struct NttTileMap {
    uint32_t q, k, L, hs_bits, block;
    uint32_t top_shift;          // != 0: L = q = 0 pass whose hs_bits passenger bits are the TOP index bits
    __device__ __forceinline__ uint32_t operator()(uint32_t e) const {
        uint32_t lq = e & ((1u << q) - 1);
        uint32_t m = (e >> q) & ((1u << k) - 1);
        uint32_t hsub = e >> (q + k);
        if (top_shift) return (hsub << top_shift) | (block << k) | m;
        uint32_t outer = (block << hs_bits) | hsub;
        uint32_t lhi = outer & ((1u << (L - q)) - 1);
        uint32_t H = outer >> (L - q);
        return (H << (k + L)) | (m << L) | (lhi << q) | lq;
    }
};

// One register-radix round: R consecutive butterfly stages on groups of 2^R tile elements that differ
// in tile bits [pos_lo, pos_lo + R).  A group is loaded from shared memory once, its R * 2^(R-1)
// butterflies run in registers (2^(R-1) independent multiplications per stage keep the multiplier pipe
// busy from a single warp), and it is stored once: one shared-memory round trip and one barrier per
// R stages instead of per stage.  The stage acting on tile bit p is stage t = k - 1 - (p - q) of the
// pass (global stage s0 + t); its twiddle index is the global index of the butterfly's lower element
// modulo 2^(L + p - q).  DIF walks the bits downwards, DIT upwards.
template <class Fr, bool DIT, int R>
__device__ __forceinline__ void ntt_round(uint4* slo, uint4* shi, const Fr* __restrict__ tw, size_t N,
                                          const NttTileMap& map, uint32_t s0, uint32_t tile, uint32_t pos_lo) {
    constexpr int GS = 1 << R;
    const uint32_t ngroups = tile >> R;
    for (uint32_t g = threadIdx.x; g < ngroups; g += blockDim.x) {
        const uint32_t base = ((g >> pos_lo) << (pos_lo + R)) | (g & ((1u << pos_lo) - 1));
        Fr x[GS];
#pragma unroll
        for (int c = 0; c < GS; c++) x[c] = sm_load<Fr>(slo, shi, base + ((uint32_t)c << pos_lo));
        const uint32_t ibase = map(base);
        const uint32_t g_lo = map.L + (pos_lo - map.q);         // global bit of tile bit pos_lo
#pragma unroll
        for (int aa = 0; aa < R; aa++) {
            const int a = DIT ? aa : (R - 1 - aa);              // bit within the group
            const uint32_t gb = g_lo + a;                       // global bit toggled by this stage
            const uint32_t t = map.k - 1 - (pos_lo + a - map.q);
            const size_t off_s = N - (N >> (s0 + t));
            const uint32_t jb = ibase & ((1u << gb) - 1);
#pragma unroll
            for (int lowb = 0; lowb < (1 << a); lowb++) {
                const Fr w = tw[off_s + jb + ((uint32_t)lowb << g_lo)];
#pragma unroll
                for (int hi = 0; hi < (1 << (R - 1 - a)); hi++) {
                    const int c0 = (hi << (a + 1)) | lowb, c1 = c0 | (1 << a);
                    if (DIT) {
                        Fr v = x[c1] * w;
                        x[c1] = x[c0] - v;
                        x[c0] = x[c0] + v;
                    } else {
                        Fr u = x[c0];
                        x[c0] = u + x[c1];
                        x[c1] = (u - x[c1]) * w;
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < GS; c++) sm_store(slo, shi, base + ((uint32_t)c << pos_lo), x[c]);
    }
}

// One pass.  grid = (N >> T, K).  in/out may alias when no permutation flag is set.
template <class Fr, bool DIT>
__global__ void __launch_bounds__(NTT_TPB, 2)
k_ntt_pass(const Fr* in_all, Fr* out_all, const Fr* __restrict__ tw, uint32_t log_n,
           uint32_t s0, uint32_t k, uint32_t q, uint32_t T, const Fr* __restrict__ pre_scale,
           const Fr* __restrict__ post_scale, const Fr* __restrict__ post_const, uint32_t flags) {
    extern __shared__ uint4 smem[];
    uint4* slo = smem;
    uint4* shi = smem + (1u << T);
    const size_t N = size_t(1) << log_n;
    const Fr* in = in_all + (size_t)blockIdx.y * N;
    Fr* out = out_all + (size_t)blockIdx.y * N;
    const uint32_t L = log_n - s0 - k;
    const uint32_t tile = 1u << T;
    const uint32_t hs_bits = T - k - q;
    const bool top = (flags & NTT_TOP_PASS) != 0;
    const NttTileMap gidx{q, k, L, hs_bits, blockIdx.x, top ? log_n - hs_bits : 0u};

    // top-passenger tiles enumerate the passenger bits fastest, so that consecutive threads touch the
    // consecutive (bit-reversed) addresses of one 128-byte run
    auto tile_elem = [&](uint32_t ee) -> uint32_t {
        return top ? (((ee & ((1u << hs_bits) - 1)) << k) | (ee >> hs_bits)) : ee;
    };
    for (uint32_t ee = threadIdx.x; ee < tile; ee += blockDim.x) {
        const uint32_t e = tile_elem(ee);
        uint32_t i = gidx(e);
        uint32_t src = (flags & NTT_PERM_IN) ? bitrev(i, log_n) : i;
        Fr x = in[src];
        if (pre_scale) x = x * pre_scale[src];
        sm_store(slo, shi, e, x);
    }
    __syncthreads();

    // the k stages act on tile bits [q, q + k): DIF from the top bit down, DIT from the bottom bit up,
    // three bits (one radix-8 round) at a time
    const uint32_t rmax = ((flags >> NTT_RMAX_SHIFT) & 3u) ? ((flags >> NTT_RMAX_SHIFT) & 3u) : 3u;
    uint32_t done = 0;
    while (done < k) {
        const uint32_t R = (k - done >= rmax) ? rmax : (k - done);
        const uint32_t pos_lo = DIT ? (q + done) : (q + k - done - R);
        if (R == 3) ntt_round<Fr, DIT, 3>(slo, shi, tw, N, gidx, s0, tile, pos_lo);
        else if (R == 2) ntt_round<Fr, DIT, 2>(slo, shi, tw, N, gidx, s0, tile, pos_lo);
        else ntt_round<Fr, DIT, 1>(slo, shi, tw, N, gidx, s0, tile, pos_lo);
        __syncthreads();
        done += R;
    }

    Fr pc;
    if (post_const) pc = *post_const;
    for (uint32_t ee = threadIdx.x; ee < tile; ee += blockDim.x) {
        const uint32_t e = tile_elem(ee);
        uint32_t i = gidx(e);
        uint32_t dst = (flags & NTT_PERM_OUT) ? bitrev(i, log_n) : i;
        Fr x = sm_load<Fr>(slo, shi, e);
        if (post_scale) x = x * post_scale[dst];
        if (post_const) x = x * pc;
        out[dst] = x;
    }
}

// ---- table construction (one-off per (context, log_n)) ------------------------------------------
// pw[b] = base^(2^b), b < nb  (single thread)
template <class Fr>
__global__ void k_pow2_table(Fr base, Fr* pw, uint32_t nb) {
    if (threadIdx.x || blockIdx.x) return;
    Fr x = base;
    for (uint32_t b = 0; b < nb; b++) {
        pw[b] = x;
        x = x.sqr();
    }
}

template <class Fr>
__device__ __forceinline__ Fr pow_from_table(const Fr* pw, uint32_t e) {
    Fr acc = Fr::one();
    for (uint32_t b = 0; e; b++, e >>= 1)
        if (e & 1) acc = acc * pw[b];
    return acc;
}

// out[i] = mult * base^(perm(i) * stride), i < n ; perm = bitrev over `bits` when bits > 0
template <class Fr>
__global__ void k_pow_fill(const Fr* pw, Fr* out, uint32_t n, uint32_t stride_log, uint32_t br_bits,
                           const Fr* mult) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t e = br_bits ? bitrev(i, br_bits) : i;
    // exponent e << stride_log can exceed 32 bits only if stride_log + log2(n) > 32; callers keep it <= 32
    uint64_t ee = (uint64_t)e << stride_log;
    Fr acc = Fr::one();
    for (uint32_t b = 0; ee; b++, ee >>= 1)
        if (ee & 1) acc = acc * pw[b];
    if (mult) acc = acc * (*mult);
    out[i] = acc;
}

// scalars: [0] = w_N, [1] = w_N^-1, [2] = g, [3] = g^-1, [4] = 1/N, [5] = 1/(g^N - 1)
template <class Curve>
__global__ void k_domain_scalars(typename Curve::Fr* sc, uint32_t log_n) {
    using Fr = typename Curve::Fr;
    if (threadIdx.x || blockIdx.x) return;
    Fr w = FrConsts<Curve>::root();
    for (uint32_t i = log_n; i < (uint32_t)Curve::TWO_ADICITY; i++) w = w.sqr();
    Fr g = FrConsts<Curve>::gen();
    sc[0] = w;
    sc[1] = w.inverse();
    sc[2] = g;
    sc[3] = g.inverse();
    // N in Montgomery form: one() doubled log_n times
    Fr nn = Fr::one();
    for (uint32_t i = 0; i < log_n; i++) nn = nn.dbl();
    sc[4] = nn.inverse();
    Fr gN = g;
    for (uint32_t i = 0; i < log_n; i++) gN = gN.sqr();
    sc[5] = (gN - Fr::one()).inverse();
}

template <class Curve>
std::shared_ptr<NttDomain> ntt_domain(hk_ctx* ctx, uint32_t log_n) {
    using Fr = typename Curve::Fr;
    // One builder at a time; forks of a context share the tables (read-only once built).
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    auto it = ctx->sh->domains.find(log_n);
    if (it != ctx->sh->domains.end()) return it->second;
    HK_REQUIRE(log_n <= (uint32_t)Curve::TWO_ADICITY, HK_ERR_UNSUPPORTED,
               "ntt: log_n exceeds the two-adicity of Fr (PolynomialDegreeTooLarge)");
    HK_REQUIRE(log_n <= 30, HK_ERR_UNSUPPORTED, "ntt: log_n > 30 not supported");
    auto d = std::make_shared<NttDomain>();
    d->log_n = log_n;
    const size_t N = size_t(1) << log_n;
    // layout: 6 tables of N + scalars(8) + pow2 tables 4 x 40
    Fr* base = d->buf.as<Fr>(6 * N + 8 + 4 * 40);
    Fr* tw_fwd = base;
    Fr* tw_inv = base + N;
    Fr* coset_nat = base + 2 * N;
    Fr* coset_inv_nat = base + 3 * N;
    Fr* coset_br = base + 4 * N;
    Fr* coset_inv_br = base + 5 * N;
    Fr* sc = base + 6 * N;
    Fr* pw_w = sc + 8;
    Fr* pw_wi = pw_w + 40;
    Fr* pw_g = pw_wi + 40;
    Fr* pw_gi = pw_g + 40;
    cudaStream_t st = ctx->stream;
    k_domain_scalars<Curve><<<1, 1, 0, st>>>(sc, log_n);
    HK_LAUNCH_CHECK(ctx);
    // the pow2 kernels take their base by value from device memory -> read it back (tiny, one-off)
    Fr h[6];
    HK_CUDA(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
    HK_CUDA(cudaStreamSynchronize(st));
    uint32_t nb = log_n + 1;
    k_pow2_table<Fr><<<1, 1, 0, st>>>(h[0], pw_w, nb);
    k_pow2_table<Fr><<<1, 1, 0, st>>>(h[1], pw_wi, nb);
    k_pow2_table<Fr><<<1, 1, 0, st>>>(h[2], pw_g, nb);
    k_pow2_table<Fr><<<1, 1, 0, st>>>(h[3], pw_gi, nb);
    ctx->launches += 4;
    HK_CUDA(cudaGetLastError());
    for (uint32_t s = 0; s < log_n; s++) {
        uint32_t hs = (uint32_t)(N >> (s + 1));
        size_t off = N - (N >> s);
        k_pow_fill<Fr><<<ceil_div(hs, 256), 256, 0, st>>>(pw_w, tw_fwd + off, hs, s, 0, nullptr);
        k_pow_fill<Fr><<<ceil_div(hs, 256), 256, 0, st>>>(pw_wi, tw_inv + off, hs, s, 0, nullptr);
        ctx->launches += 2;
    }
    HK_CUDA(cudaGetLastError());
    uint32_t n32 = (uint32_t)N;
    k_pow_fill<Fr><<<ceil_div(N, 256), 256, 0, st>>>(pw_g, coset_nat, n32, 0, 0, nullptr);
    k_pow_fill<Fr><<<ceil_div(N, 256), 256, 0, st>>>(pw_gi, coset_inv_nat, n32, 0, 0, sc + 4);
    k_pow_fill<Fr><<<ceil_div(N, 256), 256, 0, st>>>(pw_g, coset_br, n32, 0, log_n, sc + 4);
    k_pow_fill<Fr><<<ceil_div(N, 256), 256, 0, st>>>(pw_gi, coset_inv_br, n32, 0, log_n, sc + 4);
    ctx->launches += 4;
    HK_CUDA(cudaGetLastError());
    d->tw_fwd = tw_fwd;
    d->tw_inv = tw_inv;
    d->coset_nat = coset_nat;
    d->coset_inv_nat = coset_inv_nat;
    d->coset_br = coset_br;
    d->coset_inv_br = coset_inv_br;
    d->n_inv = sc + 4;
    d->zh_inv = sc + 5;
    HK_CUDA(cudaStreamSynchronize(st));   // other streams may use the tables right away
    ctx->sh->domains[log_n] = d;
    return d;
}

// Launch all passes of one transform.
//   dit = false: DIF (natural -> bit-reversed), passes in plan order
//   dit = true : DIT (bit-reversed -> natural), passes in reverse plan order
// src -> dst; `tmp` is used when a permutation forces out-of-place passes.
template <class Curve>
void ntt_passes(hk_ctx* ctx, const NttDomain& d, bool dit, bool inverse_root, const typename Curve::Fr* src,
                typename Curve::Fr* dst, typename Curve::Fr* tmp, size_t K, bool perm_in, bool perm_out,
                const typename Curve::Fr* pre_scale, const typename Curve::Fr* post_scale,
                const typename Curve::Fr* post_const, bool src_clobber_ok = false) {
    using Fr = typename Curve::Fr;
    const uint32_t log_n = d.log_n;
    // the bit reversal sits on the L = 0 pass only for DIT + perm_in (first pass) and DIF + perm_out (last pass)
    const bool top_pass = ((dit && perm_in) || (!dit && perm_out)) && log_n >= NTT_TOP_MIN && log_n > NTT_TMAX;
    const uint32_t tbits = ntt_tile_bits(log_n, K, ctx->sm_count);
    auto plan = ntt_plan(log_n, top_pass, tbits);
    const unsigned tpb = NTT_TPB;
    const uint32_t rflag = tbits >= NTT_TMAX ? 0u : ((tbits - 8u) << NTT_RMAX_SHIFT);
    const Fr* tw = reinterpret_cast<const Fr*>(inverse_root ? d.tw_inv : d.tw_fwd);
    const size_t np = plan.size();
    if (dit) {
        HK_CUDA(cudaFuncSetAttribute(k_ntt_pass<Fr, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        HK_CUDA(cudaFuncSetAttribute(k_ntt_pass<Fr, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    } else {
        HK_CUDA(cudaFuncSetAttribute(k_ntt_pass<Fr, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        HK_CUDA(cudaFuncSetAttribute(k_ntt_pass<Fr, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    // Buffer routing.  A pass carrying a permutation must run out of place; any other pass may
    // run in place.  `src` is only overwritten when src == dst or the caller allows it.
    const Fr* cur = src;
    for (size_t pi = 0; pi < np; pi++) {
        const NttPass& ps = dit ? plan[np - 1 - pi] : plan[pi];
        bool first = (pi == 0), last = (pi + 1 == np);
        uint32_t flags = 0;
        if (first && perm_in) flags |= NTT_PERM_IN;
        if (last && perm_out) flags |= NTT_PERM_OUT;
        if (top_pass && ps.s0 + ps.k == log_n) flags |= NTT_TOP_PASS;     // the L = 0 pass
        flags |= rflag;
        const bool oop = (flags & (NTT_PERM_IN | NTT_PERM_OUT)) != 0;
        Fr* to;
        if (last) {
            to = dst;
            if (oop && cur == dst) to = tmp;          // copied back below
        } else if (oop) {
            to = (cur == tmp) ? dst : tmp;
        } else if (cur == src && src != dst && !src_clobber_ok) {
            to = dst;
        } else {
            to = const_cast<Fr*>(cur);
        }
        dim3 grid((unsigned)((size_t(1) << log_n) >> ps.T), (unsigned)K);
        size_t smem = (size_t(2) << ps.T) * sizeof(uint4);
        const Fr* pre = first ? pre_scale : nullptr;
        const Fr* post = last ? post_scale : nullptr;
        const Fr* pc = last ? post_const : nullptr;
        int pidx = prof_begin(ctx, HK_PROF_NTT_PASS, ctx->qstream());
        if (dit)
            k_ntt_pass<Fr, true><<<grid, tpb, smem, ctx->qstream()>>>(cur, to, tw, log_n, ps.s0, ps.k, ps.q, ps.T, pre,
                                                                   post, pc, flags);
        else
            k_ntt_pass<Fr, false><<<grid, tpb, smem, ctx->qstream()>>>(cur, to, tw, log_n, ps.s0, ps.k, ps.q, ps.T, pre,
                                                                    post, pc, flags);
        HK_LAUNCH_CHECK(ctx);
        prof_end(ctx, pidx, (uint64_t)K << log_n, nullptr, ctx->qstream());
        cur = to;
    }
    if (cur != dst) {
        HK_CUDA(cudaMemcpyAsync(dst, cur, (K << log_n) * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->qstream()));
    }
}

template <class Curve>
void ntt_run(hk_ctx* ctx, uint64_t* d_data, uint32_t log_n, size_t K, uint32_t dir, uint32_t domain) {
    using Fr = typename Curve::Fr;
    if (K == 0) return;
    Fr* data = reinterpret_cast<Fr*>(d_data);
    if (log_n == 0) {
        return;   // size-1 transform is the identity in both directions (offset^0 = 1, 1/N = 1)
    }
    auto dom = ntt_domain<Curve>(ctx, log_n);
    const size_t N = size_t(1) << log_n;
    Fr* tmp = ctx->ws_ntt_tmp.as<Fr>(N * K);
    const bool coset = (domain == HK_DOMAIN_COSET_GENERATOR);
    if (dir == HK_NTT_FORWARD) {
        // DIT(w) with bit-reversed loads; coset: x_j *= g^j at load (index = source position)
        ntt_passes<Curve>(ctx, *dom, true, false, data, data, tmp, K, true, false,
                          coset ? reinterpret_cast<const Fr*>(dom->coset_nat) : nullptr, nullptr, nullptr);
    } else {
        // DIT(w^-1) with bit-reversed loads, then * 1/N (plain) or * g^-k / N (coset) at the store
        ntt_passes<Curve>(ctx, *dom, true, true, data, data, tmp, K, true, false, nullptr,
                          coset ? reinterpret_cast<const Fr*>(dom->coset_inv_nat) : nullptr,
                          coset ? nullptr : reinterpret_cast<const Fr*>(dom->n_inv));
    }
}

}  // namespace hk
