// extern "C" surface of libhekaton_b200.so (declared in include/hekaton_b200.h).
// Host pointers are staged through context-owned device buffers; device pointers are used in
// place.  There is no CPU compute path anywhere in this file.
#include <chrono>

#include "common.cuh"
#include "prove.cuh"
#include "qap.cuh"

using namespace hk;

#define HK_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

template <class Fn1, class Fn2>
void by_curve(hk_ctx* ctx, Fn1&& bls, Fn2&& bn) {
    if (ctx->curve == HK_CURVE_BLS12_381) bls();
    else bn();
}

size_t fq_bytes(const hk_ctx* ctx) { return ctx->curve == HK_CURVE_BLS12_381 ? 48 : 32; }
size_t point_bytes(const hk_ctx* ctx, int group) { return fq_bytes(ctx) * 2 * (group == 2 ? 2 : 1); }
size_t jac_bytes(const hk_ctx* ctx, int group) { return fq_bytes(ctx) * 3 * (group == 2 ? 2 : 1); }

// stage `bytes` from host (via pinned memory) or alias a device pointer
const void* to_device(hk_ctx* ctx, DevBuf& buf, const void* p, size_t bytes, uint32_t memspace) {
    if (memspace == HK_DEVICE || bytes == 0) return p;
    void* d = buf.ensure(bytes);
    HK_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return d;
}

// device pointers handed to kernels that use 128-bit vector accesses
void require_aligned16(const void* p, const char* what) {
    HK_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, HK_ERR_INVALID,
               std::string(what) + ": device pointers must be 16-byte aligned");
}

uint64_t new_handle(hk_ctx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    return ctx->sh->next_handle++;
}

void prof_clear(hk_ctx* ctx) {
    for (auto& r : ctx->prof) {
        if (r.e0) cudaEventDestroy(r.e0);
        if (r.e1) cudaEventDestroy(r.e1);
    }
    ctx->prof.clear();
    ctx->prof_pinned_used = 0;
}


std::shared_ptr<BaseSet> find_bases_any(hk_ctx* ctx, uint64_t h) {
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    auto it = ctx->sh->bases.find(h);
    HK_REQUIRE(it != ctx->sh->bases.end(), HK_ERR_HANDLE, "unknown base-set handle");
    return it->second;
}
std::shared_ptr<BaseSet> find_bases(hk_ctx* ctx, uint64_t h, int group) {
    auto bs = find_bases_any(ctx, h);
    HK_REQUIRE(bs->group == group, HK_ERR_INVALID, "base set belongs to the other group");
    return bs;
}
void register_bases(hk_ctx* ctx, std::shared_ptr<BaseSet> bs, uint64_t* handle) {
    uint64_t h = new_handle(ctx);
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    ctx->sh->bases[h] = std::move(bs);
    *handle = h;
}
std::shared_ptr<PkDev> find_pk(hk_ctx* ctx, uint64_t h) {
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    auto it = ctx->sh->pks.find(h);
    HK_REQUIRE(it != ctx->sh->pks.end(), HK_ERR_HANDLE, "unknown pk handle");
    return it->second;
}
std::shared_ptr<R1csDev> find_r1cs(hk_ctx* ctx, uint64_t h) {
    std::lock_guard<std::mutex> lk(ctx->sh->mu);
    auto it = ctx->sh->r1cs.find(h);
    HK_REQUIRE(it != ctx->sh->r1cs.end(), HK_ERR_HANDLE, "unknown r1cs handle");
    return it->second;
}

void bases_upload(hk_ctx* ctx, int group, const void* pts, const uint8_t* inf, size_t n, uint32_t memspace,
                  uint64_t* handle) {
    HK_REQUIRE(handle, HK_ERR_INVALID, "handle is NULL");
    HK_REQUIRE(pts || n == 0, HK_ERR_INVALID, "points pointer is NULL");
    auto bs = std::make_shared<BaseSet>();
    bs->group = group;
    bs->n = n;
    bs->point_bytes = point_bytes(ctx, group);
    size_t bytes = n * bs->point_bytes;
    cudaError_t e = cudaMalloc(&bs->d, bytes + 256);
    if (e != cudaSuccess) throw HkError{HK_ERR_NOMEM, std::string("cudaMalloc bases: ") + cudaGetErrorString(e)};
    if (bytes)
        HK_CUDA(cudaMemcpyAsync(bs->d, pts, bytes, memspace == HK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                ctx->stream));
    if (inf && n) {
        const uint8_t* d_inf = (const uint8_t*)to_device(ctx, ctx->ws_io[3], inf, n, memspace);
        by_curve(ctx, [&] { apply_infinity<Bls12381>(ctx, *bs, d_inf); }, [&] { apply_infinity<Bn254>(ctx, *bs, d_inf); });
    }
    HK_CUDA(cudaStreamSynchronize(ctx->stream));
    register_bases(ctx, bs, handle);
}

void bases_from_scalars(hk_ctx* ctx, int group, const void* gen, const uint64_t* scalars, size_t n, uint32_t repr,
                        uint32_t memspace, uint64_t* handle) {
    HK_REQUIRE(handle && gen, HK_ERR_INVALID, "NULL argument");
    auto bs = std::make_shared<BaseSet>();
    bs->group = group;
    bs->n = n;
    bs->point_bytes = point_bytes(ctx, group);
    cudaError_t e = cudaMalloc(&bs->d, n * bs->point_bytes + 256);
    if (e != cudaSuccess) throw HkError{HK_ERR_NOMEM, std::string("cudaMalloc bases: ") + cudaGetErrorString(e)};
    const void* d_gen = to_device(ctx, ctx->ws_io[2], gen, bs->point_bytes, HK_HOST);
    const Scalar256* d_sc = (const Scalar256*)to_device(ctx, ctx->ws_io[0], scalars, n * 32, memspace);
    bool mont = repr == HK_SCALAR_MONT;
    if (group == 1)
        by_curve(ctx, [&] { fixed_base_mul<Bls12381, 1>(ctx, d_gen, d_sc, n, mont, bs->d); },
                 [&] { fixed_base_mul<Bn254, 1>(ctx, d_gen, d_sc, n, mont, bs->d); });
    else
        by_curve(ctx, [&] { fixed_base_mul<Bls12381, 2>(ctx, d_gen, d_sc, n, mont, bs->d); },
                 [&] { fixed_base_mul<Bn254, 2>(ctx, d_gen, d_sc, n, mont, bs->d); });
    HK_CUDA(cudaStreamSynchronize(ctx->stream));
    register_bases(ctx, bs, handle);
}

void msm_entry(hk_ctx* ctx, int group, uint64_t bases, size_t base_offset, const uint64_t* scalars, size_t n, size_t K,
               uint32_t repr, uint32_t memspace, void* out) {
    HK_REQUIRE(out, HK_ERR_INVALID, "out is NULL");
    HK_REQUIRE(scalars || n == 0, HK_ERR_INVALID, "scalars is NULL");
    HK_REQUIRE(repr == HK_SCALAR_BIGINT || repr == HK_SCALAR_MONT, HK_ERR_INVALID, "bad scalar_repr");
    std::shared_ptr<BaseSet> bsp = find_bases(ctx, bases, group);
    BaseSet& bs = *bsp;
    // msm_unchecked / msm_bigint zip to the shorter of (bases, scalars)
    size_t avail = base_offset < bs.n ? bs.n - base_offset : 0;
    size_t len = n < avail ? n : avail;
    const size_t jb = jac_bytes(ctx, group);
    const Scalar256* d_sc;
    if (memspace == HK_DEVICE || n == 0) {
        d_sc = (const Scalar256*)scalars;
        HK_REQUIRE(len == n || K == 1, HK_ERR_UNSUPPORTED, "batched msm with truncated length needs K == 1");
    } else {
        // copy only the used prefix of each vector
        Scalar256* d = ctx->ws_io[0].as<Scalar256>(len * K + 1);
        for (size_t k = 0; k < K; k++)
            if (len)
                HK_CUDA(cudaMemcpyAsync(d + k * len, (const Scalar256*)scalars + k * n, len * 32, cudaMemcpyHostToDevice,
                                        ctx->stream));
        d_sc = d;
    }
    if (memspace == HK_DEVICE) {
        require_aligned16(scalars, "hk_msm scalars");
        require_aligned16(out, "hk_msm out");
    }
    void* d_out = (memspace == HK_DEVICE) ? out : ctx->ws_out.ensure(jb * K + 64);
    const MsmBases bv = msm_view(bs);
    const size_t off = len ? base_offset : 0;
    bool mont = repr == HK_SCALAR_MONT;
    if (group == 1)
        by_curve(ctx, [&] { msm_run<Bls12381, 1>(ctx, bv, off, len, d_sc, K, mont, d_out); },
                 [&] { msm_run<Bn254, 1>(ctx, bv, off, len, d_sc, K, mont, d_out); });
    else
        by_curve(ctx, [&] { msm_run<Bls12381, 2>(ctx, bv, off, len, d_sc, K, mont, d_out); },
                 [&] { msm_run<Bn254, 2>(ctx, bv, off, len, d_sc, K, mont, d_out); });
    if (memspace != HK_DEVICE) {
        HK_CUDA(cudaMemcpyAsync(out, d_out, jb * K, cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
}

void into_affine_entry(hk_ctx* ctx, int group, const void* jac, size_t n, void* out_xy, uint8_t* inf_out) {
    HK_REQUIRE(jac && out_xy, HK_ERR_INVALID, "NULL argument");
    if (n == 0) return;
    const size_t jb = jac_bytes(ctx, group), pb = point_bytes(ctx, group);
    const void* d_j = to_device(ctx, ctx->ws_io[0], jac, jb * n, HK_HOST);
    char* d_o = (char*)ctx->ws_io[1].ensure(pb * n + n + 64);
    uint8_t* d_inf = (uint8_t*)(d_o + pb * n);
    if (group == 1)
        by_curve(ctx, [&] { jac_to_affine<Bls12381, 1>(ctx, d_j, d_o, d_inf, n); },
                 [&] { jac_to_affine<Bn254, 1>(ctx, d_j, d_o, d_inf, n); });
    else
        by_curve(ctx, [&] { jac_to_affine<Bls12381, 2>(ctx, d_j, d_o, d_inf, n); },
                 [&] { jac_to_affine<Bn254, 2>(ctx, d_j, d_o, d_inf, n); });
    HK_CUDA(cudaMemcpyAsync(out_xy, d_o, pb * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (inf_out) HK_CUDA(cudaMemcpyAsync(inf_out, d_inf, n, cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- element-wise field ops (test hook) ---------------------------------------------------------
template <class F>
__global__ void k_field_op(const F* a, const F* b, F* out, size_t n, uint32_t op) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i];
    F r;
    switch (op) {
        case HK_FOP_MUL: r = x * b[i]; break;
        case HK_FOP_ADD: r = x + b[i]; break;
        case HK_FOP_SUB: r = x - b[i]; break;
        case HK_FOP_INV: r = x.inverse(); break;
        case HK_FOP_TO_MONT: r = x.to_mont(); break;
        case HK_FOP_INV_GCD: r = x.inverse_fast(); break;
        case HK_FOP_SQR: r = x.sqr(); break;
        case HK_FOP_MUL_KARA: r = x.mul_kara(b[i]); break;
        case HK_FOP_MULSUB2: r = F::mul_sub_mul_fused(x, b[i], b[i], x + b[i]); break;
        default: r = x.from_mont(); break;
    }
    out[i] = r;
}

template <class F>
void field_op_run(hk_ctx* ctx, uint32_t op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    size_t bytes = n * sizeof(F);
    const F* da = (const F*)to_device(ctx, ctx->ws_io[0], a, bytes, HK_HOST);
    const F* db = b ? (const F*)to_device(ctx, ctx->ws_io[1], b, bytes, HK_HOST) : da;
    F* dout = ctx->ws_io[2].as<F>(n + 1);
    k_field_op<F><<<ceil_div(n, 128), 128, 0, ctx->stream>>>(da, db, dout, n, op);
    HK_LAUNCH_CHECK(ctx);
    HK_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- throughput probes ---------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) k_bench_mul(F* out, uint32_t iters) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    F x = F::zero(), y = F::zero();
    x.r.v[0] = (i & 0xffff) + 2;   // small values are valid (non-canonical-looking but < p) Montgomery residues
    y.r.v[0] = (i & 0xfff) + 3;
    // two independent chains
    for (uint32_t k = 0; k < iters; k += 2) {
        x = x * x;
        y = y * y;
    }
    out[i] = x + y;
}

__global__ void __launch_bounds__(256) k_bench_imad_wide(uint64_t* out, uint32_t iters) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = tid * 2654435761u + 1, b = tid ^ 0x9e3779b9u;
    uint64_t c0 = tid, c1 = tid + 1, c2 = tid + 2, c3 = tid + 3, c4 = tid + 4, c5 = tid + 5, c6 = tid + 6, c7 = tid + 7;
    for (uint32_t k = 0; k < iters; k++) {
        // 8 independent 4-long carry chains worth of wide multiply-accumulates = 32 per iteration
        asm volatile(
            "{\n\t.reg .u64 t0,t1,t2,t3,t4,t5,t6,t7;\n\t"
            "mul.wide.u32 t0, %8, %9;\n\t mul.wide.u32 t1, %9, %9;\n\t mul.wide.u32 t2, %8, %8;\n\t mul.wide.u32 t3, %9, %8;\n\t"
            "add.cc.u64 %0, %0, t0;\n\t addc.cc.u64 %1, %1, t1;\n\t addc.cc.u64 %2, %2, t2;\n\t addc.u64 %3, %3, t3;\n\t"
            "mul.wide.u32 t4, %8, %9;\n\t mul.wide.u32 t5, %9, %9;\n\t mul.wide.u32 t6, %8, %8;\n\t mul.wide.u32 t7, %9, %8;\n\t"
            "add.cc.u64 %4, %4, t4;\n\t addc.cc.u64 %5, %5, t5;\n\t addc.cc.u64 %6, %6, t6;\n\t addc.u64 %7, %7, t7;\n\t"
            "mul.wide.u32 t0, %8, %9;\n\t mul.wide.u32 t1, %9, %9;\n\t mul.wide.u32 t2, %8, %8;\n\t mul.wide.u32 t3, %9, %8;\n\t"
            "add.cc.u64 %1, %1, t0;\n\t addc.cc.u64 %2, %2, t1;\n\t addc.cc.u64 %3, %3, t2;\n\t addc.u64 %0, %0, t3;\n\t"
            "mul.wide.u32 t4, %8, %9;\n\t mul.wide.u32 t5, %9, %9;\n\t mul.wide.u32 t6, %8, %8;\n\t mul.wide.u32 t7, %9, %8;\n\t"
            "add.cc.u64 %5, %5, t4;\n\t addc.cc.u64 %6, %6, t5;\n\t addc.cc.u64 %7, %7, t6;\n\t addc.u64 %4, %4, t7;\n\t"
            "mul.wide.u32 t0, %8, %9;\n\t mul.wide.u32 t1, %9, %9;\n\t mul.wide.u32 t2, %8, %8;\n\t mul.wide.u32 t3, %9, %8;\n\t"
            "add.cc.u64 %2, %2, t0;\n\t addc.cc.u64 %3, %3, t1;\n\t addc.cc.u64 %0, %0, t2;\n\t addc.u64 %1, %1, t3;\n\t"
            "mul.wide.u32 t4, %8, %9;\n\t mul.wide.u32 t5, %9, %9;\n\t mul.wide.u32 t6, %8, %8;\n\t mul.wide.u32 t7, %9, %8;\n\t"
            "add.cc.u64 %6, %6, t4;\n\t addc.cc.u64 %7, %7, t5;\n\t addc.cc.u64 %4, %4, t6;\n\t addc.u64 %5, %5, t7;\n\t"
            "mul.wide.u32 t0, %8, %9;\n\t mul.wide.u32 t1, %9, %9;\n\t mul.wide.u32 t2, %8, %8;\n\t mul.wide.u32 t3, %9, %8;\n\t"
            "add.cc.u64 %3, %3, t0;\n\t addc.cc.u64 %0, %0, t1;\n\t addc.cc.u64 %1, %1, t2;\n\t addc.u64 %2, %2, t3;\n\t"
            "mul.wide.u32 t4, %8, %9;\n\t mul.wide.u32 t5, %9, %9;\n\t mul.wide.u32 t6, %8, %8;\n\t mul.wide.u32 t7, %9, %8;\n\t"
            "add.cc.u64 %7, %7, t4;\n\t addc.cc.u64 %4, %4, t5;\n\t addc.cc.u64 %5, %5, t6;\n\t addc.u64 %6, %6, t7;\n\t"
            "}"
            : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7)
            : "r"(a), "r"(b));
    }
    out[tid] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

template <class Fn>
float time_ms(hk_ctx* ctx, Fn&& fn) {
    cudaEvent_t e0, e1;
    HK_CUDA(cudaEventCreate(&e0));
    HK_CUDA(cudaEventCreate(&e1));
    HK_CUDA(cudaEventRecord(e0, ctx->stream));
    fn();
    HK_CUDA(cudaEventRecord(e1, ctx->stream));
    HK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    HK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

}  // namespace

// ==================================================================================================
static int32_t ctx_make_side(hk_ctx* ctx) {
    // numerically lower = higher priority; `least` is what default-priority (caller) streams have
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return HK_ERR_CUDA;
    auto ev = [](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
    for (int l = 0; l < hk_ctx::LANES; l++) {
        if (!ev(&ctx->ev_sorted[l]) || !ev(&ctx->ev_ready[l]) || !ev(&ctx->ev_accum[l]) || !ev(&ctx->ev_lane_done[l]))
            return HK_ERR_CUDA;
        if (l == 0) continue;          // lane 0 is the caller's stream
        int bulk = greatest + l < least ? greatest + l : least;
        if (cudaStreamCreateWithPriority(&ctx->lane_ctl[l], cudaStreamNonBlocking, greatest) != cudaSuccess) return HK_ERR_CUDA;
        if (cudaStreamCreateWithPriority(&ctx->lane_bulk[l], cudaStreamNonBlocking, bulk) != cudaSuccess) return HK_ERR_CUDA;
    }
    if (!ev(&ctx->ev_fork)) return HK_ERR_CUDA;
    return HK_OK;
}
static void ctx_free_side(hk_ctx* ctx) {
    for (int l = 0; l < hk_ctx::LANES; l++) {
        for (cudaStream_t st : {ctx->lane_ctl[l], ctx->lane_bulk[l]})
            if (st) {
                cudaStreamSynchronize(st);
                cudaStreamDestroy(st);
            }
        for (cudaEvent_t e : {ctx->ev_sorted[l], ctx->ev_ready[l], ctx->ev_accum[l], ctx->ev_lane_done[l]})
            if (e) cudaEventDestroy(e);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
}
HK_EXPORT int32_t hk_ctx_create(int device, uint32_t curve, void* stream, hk_ctx** out) {
    if (!out) return HK_ERR_INVALID;
    *out = nullptr;
    if (curve != HK_CURVE_BLS12_381 && curve != HK_CURVE_BN254) return HK_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return HK_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HK_ERR_NO_DEVICE;
    if (prop.major != 10) return HK_ERR_NO_DEVICE;   // built for sm_100a only; no fallback
    hk_ctx* ctx = new (std::nothrow) hk_ctx();
    if (!ctx) return HK_ERR_NOMEM;
    ctx->device = device;
    ctx->curve = curve;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->sh = std::make_shared<SharedState>();
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return HK_ERR_CUDA;
    }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return HK_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (ctx_make_side(ctx) != HK_OK) {
        ctx_free_side(ctx);
        delete ctx;
        return HK_ERR_CUDA;
    }
    *out = ctx;
    return HK_OK;
}

HK_EXPORT int32_t hk_ctx_destroy(hk_ctx* ctx) {
    if (!ctx) return HK_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    prof_clear(ctx);
    if (ctx->prof_pinned) cudaFreeHost(ctx->prof_pinned);
    ctx->sh.reset();     // shared resources are freed with the last context that references them
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx_free_side(ctx);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(ctx->ev_h2d[i]);
            cudaEventDestroy(ctx->ev_in_free[i]);
        }
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HK_OK;
}

HK_EXPORT int32_t hk_ctx_fork(hk_ctx* parent, void* stream, hk_ctx** out) {
    if (!parent || !out) return HK_ERR_INVALID;
    *out = nullptr;
    hk_ctx* ctx = new (std::nothrow) hk_ctx();
    if (!ctx) return HK_ERR_NOMEM;
    ctx->device = parent->device;
    ctx->curve = parent->curve;
    ctx->sm_count = parent->sm_count;
    ctx->sh = parent->sh;
    ctx->opt_accum_waves = parent->opt_accum_waves;
    ctx->opt_accum_smin = parent->opt_accum_smin;
    ctx->opt_fold_seg = parent->opt_fold_seg;
    ctx->opt_reduce_g = parent->opt_reduce_g;
    ctx->opt_sort_global = parent->opt_sort_global;
    ctx->opt_g2_pair = parent->opt_g2_pair;
    ctx->opt_quad_max = parent->opt_quad_max;
    ctx->opt_no_fold_resolve = parent->opt_no_fold_resolve;
    if (cudaSetDevice(ctx->device) != cudaSuccess) {
        delete ctx;
        return HK_ERR_CUDA;
    }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return HK_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (ctx_make_side(ctx) != HK_OK) {
        ctx_free_side(ctx);
        delete ctx;
        return HK_ERR_CUDA;
    }
    *out = ctx;
    return HK_OK;
}

HK_EXPORT int32_t hk_ctx_set_option(hk_ctx* ctx, uint32_t option, uint64_t value) {
    return guarded(ctx, [&] {
        switch (option) {
            case HK_OPT_ACCUM_WAVES:
                HK_REQUIRE(value <= 64, HK_ERR_INVALID, "HK_OPT_ACCUM_WAVES must be 0..64");
                ctx->opt_accum_waves = (uint32_t)value;
                break;
            case HK_OPT_ACCUM_SEG_MIN:
                HK_REQUIRE(value <= 4096, HK_ERR_INVALID, "HK_OPT_ACCUM_SEG_MIN must be 0..4096");
                ctx->opt_accum_smin = (uint32_t)value;
                break;
            case HK_OPT_FOLD_SEG:
                HK_REQUIRE(value == 0 || value == 4 || value == 8 || value == 16 || value == 32, HK_ERR_INVALID,
                           "HK_OPT_FOLD_SEG must be 0, 4, 8, 16 or 32");
                ctx->opt_fold_seg = (uint32_t)value;
                break;
            case HK_OPT_REDUCE_CHUNK:
                HK_REQUIRE(value <= 65536 && (value & (value - 1)) == 0, HK_ERR_INVALID,
                           "HK_OPT_REDUCE_CHUNK must be 0 or a power of two");
                ctx->opt_reduce_g = (uint32_t)value;
                break;
            case HK_OPT_SORT_GLOBAL_ATOMICS: ctx->opt_sort_global = value ? 1u : 0u; break;
            case HK_OPT_G2_LANE_PAIR:
                HK_REQUIRE(value <= 2, HK_ERR_INVALID, "HK_OPT_G2_LANE_PAIR must be 0, 1 (on) or 2 (off)");
                ctx->opt_g2_pair = (uint32_t)value;
                break;
            case HK_OPT_QUAD_MAX_THREADS:
                HK_REQUIRE(value <= (1u << 24), HK_ERR_INVALID, "HK_OPT_QUAD_MAX_THREADS must be 0..2^24");
                ctx->opt_quad_max = (uint32_t)value;
                break;
            case HK_OPT_NO_FOLD_RESOLVE: ctx->opt_no_fold_resolve = value ? 1u : 0u; break;
            default: throw HkError{HK_ERR_INVALID, "unknown option"};
        }
    });
}

HK_EXPORT int32_t hk_ctx_sync(hk_ctx* ctx) {
    return guarded(ctx, [&] { HK_CUDA(cudaStreamSynchronize(ctx->stream)); });
}

HK_EXPORT const char* hk_last_error_string(hk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

HK_EXPORT int32_t hk_curve_info(hk_ctx* ctx, uint32_t* fq_limbs64, uint32_t* fr_limbs64, uint32_t* fr_bits,
                                uint32_t* two_adicity) {
    if (!ctx) return HK_ERR_INVALID;
    bool bls = ctx->curve == HK_CURVE_BLS12_381;
    if (fq_limbs64) *fq_limbs64 = bls ? 6 : 4;
    if (fr_limbs64) *fr_limbs64 = 4;
    if (fr_bits) *fr_bits = bls ? 255 : 254;
    if (two_adicity) *two_adicity = bls ? 32 : 28;
    return HK_OK;
}

HK_EXPORT uint64_t hk_launch_count(hk_ctx* ctx) { return ctx ? ctx->launches : 0; }

HK_EXPORT int32_t hk_bases_upload_g1(hk_ctx* ctx, const void* p, const uint8_t* inf, size_t n, uint32_t memspace,
                                     uint64_t* handle) {
    return guarded(ctx, [&] { bases_upload(ctx, 1, p, inf, n, memspace, handle); });
}
HK_EXPORT int32_t hk_bases_upload_g2(hk_ctx* ctx, const void* p, const uint8_t* inf, size_t n, uint32_t memspace,
                                     uint64_t* handle) {
    return guarded(ctx, [&] { bases_upload(ctx, 2, p, inf, n, memspace, handle); });
}
HK_EXPORT int32_t hk_bases_free(hk_ctx* ctx, uint64_t handle) {
    return guarded(ctx, [&] {
        HK_CUDA(cudaDeviceSynchronize());      // forks may still be reading the buffers on their own streams
        std::lock_guard<std::mutex> lk(ctx->sh->mu);
        auto it = ctx->sh->bases.find(handle);
        HK_REQUIRE(it != ctx->sh->bases.end(), HK_ERR_HANDLE, "unknown base-set handle");
        ctx->sh->bases.erase(it);     // memory is released when the last user (e.g. a pk) drops it
    });
}
HK_EXPORT int32_t hk_bases_len(hk_ctx* ctx, uint64_t handle, size_t* n) {
    return guarded(ctx, [&] {
        HK_REQUIRE(n, HK_ERR_INVALID, "n is NULL");
        *n = find_bases_any(ctx, handle)->n;
    });
}
HK_EXPORT int32_t hk_bases_precompute(hk_ctx* ctx, uint64_t handle, uint32_t density_permille) {
    return guarded(ctx, [&] {
        auto bs = find_bases_any(ctx, handle);
        HK_REQUIRE(density_permille <= 1000, HK_ERR_INVALID, "density_permille must be 0..1000");
        const double dn = density_permille ? density_permille / 1000.0 : 1.0;
        if (bs->group == 1)
            by_curve(ctx, [&] { bases_precompute<Bls12381, 1>(ctx, *bs, dn); }, [&] { bases_precompute<Bn254, 1>(ctx, *bs, dn); });
        else
            by_curve(ctx, [&] { bases_precompute<Bls12381, 2>(ctx, *bs, dn); }, [&] { bases_precompute<Bn254, 2>(ctx, *bs, dn); });
    });
}
HK_EXPORT int32_t hk_bases_download(hk_ctx* ctx, uint64_t handle, size_t offset, size_t n, void* out) {
    return guarded(ctx, [&] {
        auto bsp = find_bases_any(ctx, handle);
        BaseSet& bs = *bsp;
        HK_REQUIRE(out && offset + n <= bs.n, HK_ERR_INVALID, "range outside the base set");
        if (n == 0) return;
        HK_CUDA(cudaMemcpyAsync(out, (char*)bs.d + offset * bs.point_bytes, n * bs.point_bytes, cudaMemcpyDeviceToHost,
                                ctx->stream));
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}
HK_EXPORT int32_t hk_bases_from_scalars_g1(hk_ctx* ctx, const void* gen, const uint64_t* scalars, size_t n,
                                           uint32_t repr, uint32_t memspace, uint64_t* handle) {
    return guarded(ctx, [&] { bases_from_scalars(ctx, 1, gen, scalars, n, repr, memspace, handle); });
}
HK_EXPORT int32_t hk_bases_from_scalars_g2(hk_ctx* ctx, const void* gen, const uint64_t* scalars, size_t n,
                                           uint32_t repr, uint32_t memspace, uint64_t* handle) {
    return guarded(ctx, [&] { bases_from_scalars(ctx, 2, gen, scalars, n, repr, memspace, handle); });
}

// element-wise scalar multiplication (host or device buffers in `memspace`)
static void scalar_mul_entry(hk_ctx* ctx, int group, const void* pts, const uint8_t* inf, const uint64_t* scalars, size_t n,
                             uint32_t repr, uint32_t memspace, void* out_xy, uint8_t* inf_out) {
    HK_REQUIRE(n == 0 || (pts && scalars && out_xy), HK_ERR_INVALID, "NULL argument");
    HK_REQUIRE(repr == HK_SCALAR_BIGINT || repr == HK_SCALAR_MONT, HK_ERR_INVALID, "bad scalar representation");
    if (n == 0) return;
    const size_t pb = point_bytes(ctx, group);
    const bool mont = repr == HK_SCALAR_MONT;
    if (memspace == HK_DEVICE) {
        by_curve(ctx,
                 [&] {
                     if (group == 1) scalar_mul_batch<Bls12381, 1>(ctx, pts, inf, (const Scalar256*)scalars, n, mont, out_xy, inf_out);
                     else scalar_mul_batch<Bls12381, 2>(ctx, pts, inf, (const Scalar256*)scalars, n, mont, out_xy, inf_out);
                 },
                 [&] {
                     if (group == 1) scalar_mul_batch<Bn254, 1>(ctx, pts, inf, (const Scalar256*)scalars, n, mont, out_xy, inf_out);
                     else scalar_mul_batch<Bn254, 2>(ctx, pts, inf, (const Scalar256*)scalars, n, mont, out_xy, inf_out);
                 });
        return;
    }
    const void* d_p = to_device(ctx, ctx->ws_io[0], pts, pb * n, memspace);
    const Scalar256* d_s = (const Scalar256*)to_device(ctx, ctx->ws_io[1], scalars, 32 * n, memspace);
    const uint8_t* d_i = inf ? (const uint8_t*)to_device(ctx, ctx->ws_io[3], inf, n, memspace) : nullptr;
    char* d_o = (char*)ctx->ws_io[2].ensure(pb * n + n + 64);
    uint8_t* d_io = (uint8_t*)(d_o + pb * n);
    scalar_mul_entry(ctx, group, d_p, d_i, (const uint64_t*)d_s, n, repr, HK_DEVICE, d_o, d_io);
    HK_CUDA(cudaMemcpyAsync(out_xy, d_o, pb * n, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint8_t> tmp(n);
    HK_CUDA(cudaMemcpyAsync(tmp.data(), d_io, n, cudaMemcpyDeviceToHost, ctx->stream));
    HK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (inf_out) memcpy(inf_out, tmp.data(), n);
}
HK_EXPORT int32_t hk_scalar_mul_g1(hk_ctx* ctx, const void* points, const uint8_t* infinity, const uint64_t* scalars,
                                   size_t n, uint32_t repr, uint32_t memspace, void* out_xy, uint8_t* inf_out) {
    return guarded(ctx, [&] { scalar_mul_entry(ctx, 1, points, infinity, scalars, n, repr, memspace, out_xy, inf_out); });
}
HK_EXPORT int32_t hk_scalar_mul_g2(hk_ctx* ctx, const void* points, const uint8_t* infinity, const uint64_t* scalars,
                                   size_t n, uint32_t repr, uint32_t memspace, void* out_xy, uint8_t* inf_out) {
    return guarded(ctx, [&] { scalar_mul_entry(ctx, 2, points, infinity, scalars, n, repr, memspace, out_xy, inf_out); });
}

HK_EXPORT int32_t hk_msm_g1(hk_ctx* ctx, uint64_t bases, size_t off, const uint64_t* sc, size_t n, uint32_t repr,
                            uint32_t memspace, void* out) {
    return guarded(ctx, [&] { msm_entry(ctx, 1, bases, off, sc, n, 1, repr, memspace, out); });
}
HK_EXPORT int32_t hk_msm_g2(hk_ctx* ctx, uint64_t bases, size_t off, const uint64_t* sc, size_t n, uint32_t repr,
                            uint32_t memspace, void* out) {
    return guarded(ctx, [&] { msm_entry(ctx, 2, bases, off, sc, n, 1, repr, memspace, out); });
}
HK_EXPORT int32_t hk_msm_g1_checked(hk_ctx* ctx, uint64_t bases, size_t off, size_t n_bases_expected,
                                    const uint64_t* sc, size_t n, uint32_t repr, uint32_t memspace, void* out) {
    return guarded(ctx, [&] {
        // VariableBaseMSM::msm: Err(min_len) when bases.len() != scalars.len()
        HK_REQUIRE(n_bases_expected == n, HK_ERR_LENGTH,
                   "msm: bases and scalars differ in length (min = " +
                       std::to_string(n_bases_expected < n ? n_bases_expected : n) + ")");
        auto bs = find_bases(ctx, bases, 1);
        HK_REQUIRE(off + n <= bs->n, HK_ERR_LENGTH, "msm: base range outside the base set");
        msm_entry(ctx, 1, bases, off, sc, n, 1, repr, memspace, out);
    });
}
HK_EXPORT int32_t hk_msm_g1_batch(hk_ctx* ctx, uint64_t bases, size_t off, const uint64_t* sc, size_t n, size_t K,
                                  uint32_t repr, uint32_t memspace, void* out) {
    return guarded(ctx, [&] { msm_entry(ctx, 1, bases, off, sc, n, K, repr, memspace, out); });
}
HK_EXPORT int32_t hk_msm_g2_batch(hk_ctx* ctx, uint64_t bases, size_t off, const uint64_t* sc, size_t n, size_t K,
                                  uint32_t repr, uint32_t memspace, void* out) {
    return guarded(ctx, [&] { msm_entry(ctx, 2, bases, off, sc, n, K, repr, memspace, out); });
}

HK_EXPORT int32_t hk_g1_into_affine(hk_ctx* ctx, const void* jac, size_t n, void* out_xy, uint8_t* inf_out) {
    return guarded(ctx, [&] { into_affine_entry(ctx, 1, jac, n, out_xy, inf_out); });
}
HK_EXPORT int32_t hk_g2_into_affine(hk_ctx* ctx, const void* jac, size_t n, void* out_xy, uint8_t* inf_out) {
    return guarded(ctx, [&] { into_affine_entry(ctx, 2, jac, n, out_xy, inf_out); });
}

HK_EXPORT int32_t hk_ntt_batch(hk_ctx* ctx, uint64_t* inout, uint32_t log_n, size_t K, uint32_t dir, uint32_t domain,
                               uint32_t memspace) {
    return guarded(ctx, [&] {
        HK_REQUIRE(inout || K == 0, HK_ERR_INVALID, "inout is NULL");
        HK_REQUIRE(dir == HK_NTT_FORWARD || dir == HK_NTT_INVERSE, HK_ERR_INVALID, "bad dir");
        HK_REQUIRE(domain == HK_DOMAIN_PLAIN || domain == HK_DOMAIN_COSET_GENERATOR, HK_ERR_INVALID, "bad domain");
        uint32_t adic = ctx->curve == HK_CURVE_BLS12_381 ? 32 : 28;
        HK_REQUIRE(log_n <= adic, HK_ERR_UNSUPPORTED, "ntt: log_n exceeds the two-adicity of Fr");
        if (K == 0) return;
        size_t bytes = (K << log_n) * 32;
        uint64_t* d = inout;
        if (memspace != HK_DEVICE) {
            d = ctx->ws_io[0].as<uint64_t>(bytes / 8);
            HK_CUDA(cudaMemcpyAsync(d, inout, bytes, cudaMemcpyHostToDevice, ctx->stream));
        }
        by_curve(ctx, [&] { ntt_run<Bls12381>(ctx, d, log_n, K, dir, domain); },
                 [&] { ntt_run<Bn254>(ctx, d, log_n, K, dir, domain); });
        if (memspace != HK_DEVICE) {
            HK_CUDA(cudaMemcpyAsync(inout, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    });
}
HK_EXPORT int32_t hk_ntt(hk_ctx* ctx, uint64_t* inout, uint32_t log_n, uint32_t dir, uint32_t domain, uint32_t memspace) {
    return hk_ntt_batch(ctx, inout, log_n, 1, dir, domain, memspace);
}

HK_EXPORT int32_t hk_qap_h(hk_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                           uint32_t memspace, uint64_t* h) {
    return guarded(ctx, [&] {
        HK_REQUIRE(a && b && c && h, HK_ERR_INVALID, "NULL argument");
        uint32_t adic = ctx->curve == HK_CURVE_BLS12_381 ? 32 : 28;
        HK_REQUIRE(log_n >= 1 && log_n <= adic, HK_ERR_UNSUPPORTED, "qap_h: log_n out of range");
        const size_t N = size_t(1) << log_n;
        uint64_t* abc = ctx->ws_qap[0].as<uint64_t>(4 * 3 * N);
        cudaMemcpyKind kind = memspace == HK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        HK_CUDA(cudaMemcpyAsync(abc, a, N * 32, kind, ctx->stream));
        HK_CUDA(cudaMemcpyAsync(abc + 4 * N, b, N * 32, kind, ctx->stream));
        HK_CUDA(cudaMemcpyAsync(abc + 8 * N, c, N * 32, kind, ctx->stream));
        uint64_t* dh = memspace == HK_DEVICE ? h : ctx->ws_qap[1].as<uint64_t>(4 * N);
        by_curve(ctx, [&] { qap_h_from_evals<Bls12381>(ctx, log_n, abc, dh); },
                 [&] { qap_h_from_evals<Bn254>(ctx, log_n, abc, dh); });
        if (memspace != HK_DEVICE) {
            HK_CUDA(cudaMemcpyAsync(h, dh, N * 32, cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    });
}

HK_EXPORT int32_t hk_r1cs_upload(hk_ctx* ctx, size_t n, size_t l, size_t nw, const uint32_t* a_ptr,
                                 const uint32_t* a_col, const uint64_t* a_val, const uint32_t* b_ptr,
                                 const uint32_t* b_col, const uint64_t* b_val, const uint32_t* c_ptr,
                                 const uint32_t* c_col, const uint64_t* c_val, uint64_t* handle) {
    return guarded(ctx, [&] {
        HK_REQUIRE(handle && a_ptr && b_ptr && c_ptr, HK_ERR_INVALID, "NULL argument");
        const uint32_t* ptr[3] = {a_ptr, b_ptr, c_ptr};
        const uint32_t* col[3] = {a_col, b_col, c_col};
        const uint64_t* val[3] = {a_val, b_val, c_val};
        auto m = r1cs_upload(ctx, n, l, nw, ptr, col, val);
        uint32_t adic = ctx->curve == HK_CURVE_BLS12_381 ? 32 : 28;
        HK_REQUIRE(m->log_n <= adic, HK_ERR_UNSUPPORTED, "r1cs: domain larger than 2^two_adicity");
        uint64_t h = new_handle(ctx);
        {
            std::lock_guard<std::mutex> lk(ctx->sh->mu);
            ctx->sh->r1cs[h] = m;
        }
        *handle = h;
    });
}
HK_EXPORT int32_t hk_r1cs_free(hk_ctx* ctx, uint64_t handle) {
    return guarded(ctx, [&] {
        HK_CUDA(cudaDeviceSynchronize());      // forks may still be reading the buffers on their own streams
        std::lock_guard<std::mutex> lk(ctx->sh->mu);
        auto it = ctx->sh->r1cs.find(handle);
        HK_REQUIRE(it != ctx->sh->r1cs.end(), HK_ERR_HANDLE, "unknown r1cs handle");
        ctx->sh->r1cs.erase(it);
    });
}
HK_EXPORT int32_t hk_qap_witness_map(hk_ctx* ctx, uint64_t r1cs, const uint64_t* z, uint32_t memspace, uint64_t* h) {
    return guarded(ctx, [&] {
        auto msp = find_r1cs(ctx, r1cs);
        HK_REQUIRE(z && h, HK_ERR_INVALID, "NULL argument");
        R1csDev& m = *msp;
        const size_t N = size_t(1) << m.log_n;
        const uint64_t* dz = (const uint64_t*)to_device(ctx, ctx->ws_z, z, (m.l + m.nw) * 32, memspace);
        uint64_t* abc = ctx->ws_qap[0].as<uint64_t>(4 * 3 * N);
        uint64_t* dh = memspace == HK_DEVICE ? h : ctx->ws_qap[1].as<uint64_t>(4 * N);
        by_curve(ctx,
                 [&] {
                     r1cs_evals<Bls12381>(ctx, m, dz, abc);
                     qap_h_from_evals<Bls12381>(ctx, m.log_n, abc, dh);
                 },
                 [&] {
                     r1cs_evals<Bn254>(ctx, m, dz, abc);
                     qap_h_from_evals<Bn254>(ctx, m.log_n, abc, dh);
                 });
        if (memspace != HK_DEVICE) {
            HK_CUDA(cudaMemcpyAsync(h, dh, N * 32, cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    });
}

HK_EXPORT int32_t hk_pk_create(hk_ctx* ctx, const hk_pk_desc* desc, uint64_t* pk_handle) {
    return guarded(ctx, [&] {
        HK_REQUIRE(desc && pk_handle, HK_ERR_INVALID, "NULL argument");
        HK_REQUIRE(desc->alpha_g && desc->beta_g && desc->beta_h && desc->last_delta_h, HK_ERR_INVALID,
                   "pk: vk points missing");
        std::shared_ptr<PkDev> pk;
        by_curve(ctx, [&] { pk = pk_create<Bls12381>(ctx, *desc); }, [&] { pk = pk_create<Bn254>(ctx, *desc); });
        uint64_t h = new_handle(ctx);
        {
            std::lock_guard<std::mutex> lk(ctx->sh->mu);
            ctx->sh->pks[h] = pk;
        }
        *pk_handle = h;
    });
}
HK_EXPORT int32_t hk_pk_free(hk_ctx* ctx, uint64_t pk_handle) {
    return guarded(ctx, [&] {
        HK_CUDA(cudaDeviceSynchronize());      // forks may still be reading the buffers on their own streams
        std::lock_guard<std::mutex> lk(ctx->sh->mu);
        auto it = ctx->sh->pks.find(pk_handle);
        HK_REQUIRE(it != ctx->sh->pks.end(), HK_ERR_HANDLE, "unknown pk handle");
        ctx->sh->pks.erase(it);
    });
}

// copy stream + events of the asynchronous host-buffer path, created on first use
static void async_streams(hk_ctx* ctx) {
    if (ctx->copy_stream) return;
    HK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        HK_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        HK_CUDA(cudaEventCreateWithFlags(&ctx->ev_in_free[i], cudaEventDisableTiming));
    }
}

HK_EXPORT int32_t hk_commit_batch(hk_ctx* ctx, uint64_t pk_h, uint32_t stage, size_t K, const uint64_t* z,
                                  size_t z_stride, const uint64_t* kappa, uint32_t memspace, void* com_affine,
                                  uint8_t* com_inf) {
    return guarded(ctx, [&] {
        auto pksp = find_pk(ctx, pk_h);
        HK_REQUIRE(z && kappa && com_affine, HK_ERR_INVALID, "NULL argument");
        PkDev& pk = *pksp;
        HK_REQUIRE(stage < pk.num_stages, HK_ERR_INVALID, "commit: no such stage");
        if (K == 0) return;
        const size_t need = pk.r1cs->l + pk.stage_start[stage + 1];
        HK_REQUIRE(z_stride >= need || K == 1, HK_ERR_INVALID, "commit: z_stride shorter than the committed prefix");
        const size_t pb = point_bytes(ctx, 1);
        if (memspace == HK_DEVICE) {
            HK_REQUIRE(com_inf, HK_ERR_INVALID, "com_inf is NULL");
            require_aligned16(z, "hk_commit z");
            require_aligned16(kappa, "hk_commit kappa");
            require_aligned16(com_affine, "hk_commit com_affine");
            by_curve(ctx, [&] { commit_run<Bls12381>(ctx, pk, stage, K, z, z_stride, kappa, com_affine, com_inf); },
                     [&] { commit_run<Bn254>(ctx, pk, stage, K, z, z_stride, kappa, com_affine, com_inf); });
            return;
        }
        HK_REQUIRE(memspace == HK_HOST || memspace == HK_HOST_PINNED || memspace == HK_HOST_PINNED_ASYNC, HK_ERR_INVALID,
                   "bad memspace");
        const bool async = memspace == HK_HOST_PINNED_ASYNC;
        HK_REQUIRE(!async || com_inf, HK_ERR_INVALID, "com_inf is NULL");
        // host: only the committed prefix of every assignment is copied (stride `need` on the device)
        char* dz = (char*)ctx->ws_z.ensure(K * need * 32 + K * 32 + 64);
        for (size_t k = 0; k < K; k++)
            HK_CUDA(cudaMemcpyAsync(dz + k * need * 32, (const char*)z + k * z_stride * 32, need * 32, cudaMemcpyHostToDevice,
                                    ctx->stream));
        char* dk = dz + K * need * 32;
        HK_CUDA(cudaMemcpyAsync(dk, kappa, K * 32, cudaMemcpyHostToDevice, ctx->stream));
        char* d_out = (char*)ctx->ws_io[3].ensure(K * (pb + 1) + 64);
        uint8_t* d_inf = (uint8_t*)(d_out + K * pb);
        by_curve(ctx,
                 [&] { commit_run<Bls12381>(ctx, pk, stage, K, (const uint64_t*)dz, need, (const uint64_t*)dk, d_out, d_inf); },
                 [&] { commit_run<Bn254>(ctx, pk, stage, K, (const uint64_t*)dz, need, (const uint64_t*)dk, d_out, d_inf); });
        HK_CUDA(cudaMemcpyAsync(com_affine, d_out, K * pb, cudaMemcpyDeviceToHost, ctx->stream));
        if (async) {      // stream-ordered: the next call's uploads into ws_z queue behind these kernels
            HK_CUDA(cudaMemcpyAsync(com_inf, d_inf, K, cudaMemcpyDeviceToHost, ctx->stream));
            return;
        }
        std::vector<uint8_t> tmp(K);
        HK_CUDA(cudaMemcpyAsync(tmp.data(), d_inf, K, cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (com_inf) memcpy(com_inf, tmp.data(), K);
    });
}

HK_EXPORT int32_t hk_commit(hk_ctx* ctx, uint64_t pk_h, uint32_t stage, const uint64_t* z, const uint64_t* kappa,
                            uint32_t memspace, void* com_affine, uint8_t* com_inf) {
    uint8_t scratch = 0;
    // the flag is written asynchronously for HK_DEVICE / HK_HOST_PINNED_ASYNC: a stack scratch would dangle
    if ((memspace == HK_DEVICE || memspace == HK_HOST_PINNED_ASYNC) && !com_inf) return HK_ERR_INVALID;
    return hk_commit_batch(ctx, pk_h, stage, 1, z, 0, kappa, memspace, com_affine, com_inf ? com_inf : &scratch);
}

HK_EXPORT int32_t hk_prove_batch(hk_ctx* ctx, uint64_t pk_h, size_t K, const uint64_t* z, const uint64_t* r,
                                 const uint64_t* s, const uint64_t* kappas, uint32_t memspace, void* a_affine,
                                 void* b_affine, void* c_affine, uint8_t* inf) {
    return guarded(ctx, [&] {
        auto pksp = find_pk(ctx, pk_h);
        PkDev& pk = *pksp;
        if (K == 0) return;
        HK_REQUIRE(z && r && s && a_affine && b_affine && c_affine, HK_ERR_INVALID, "NULL argument");
        HK_REQUIRE(kappas || pk.num_stages == 1, HK_ERR_INVALID, "kappas is NULL");
        const size_t p1 = point_bytes(ctx, 1), p2 = point_bytes(ctx, 2);
        const uint32_t nk = pk.num_stages - 1;
        if (memspace == HK_DEVICE) {
            HK_REQUIRE(inf, HK_ERR_INVALID, "inf is NULL");
            require_aligned16(z, "hk_prove z");
            require_aligned16(r, "hk_prove r");
            require_aligned16(s, "hk_prove s");
            if (kappas) require_aligned16(kappas, "hk_prove kappas");
            require_aligned16(a_affine, "hk_prove a");
            require_aligned16(b_affine, "hk_prove b");
            require_aligned16(c_affine, "hk_prove c");
            by_curve(ctx, [&] { prove_run<Bls12381>(ctx, pk, K, z, r, s, kappas, a_affine, b_affine, c_affine, inf); },
                     [&] { prove_run<Bn254>(ctx, pk, K, z, r, s, kappas, a_affine, b_affine, c_affine, inf); });
            return;
        }
        if (memspace == HK_HOST_PINNED_ASYNC) {
            HK_REQUIRE(inf, HK_ERR_INVALID, "inf is NULL");
            const size_t zb = K * pk.m * 32, rb = K * 32, kb = K * (size_t)nk * 32;
            async_streams(ctx);
            const int slot = (int)(ctx->async_seq++ & 1);
            char* dz = (char*)ctx->ws_zin[slot].ensure(zb + 2 * rb + kb + 64);
            // upload on the copy stream once the previous user of this staging buffer has finished
            HK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_in_free[slot], 0));
            HK_CUDA(cudaMemcpyAsync(dz, z, zb, cudaMemcpyHostToDevice, ctx->copy_stream));
            HK_CUDA(cudaMemcpyAsync(dz + zb, r, rb, cudaMemcpyHostToDevice, ctx->copy_stream));
            HK_CUDA(cudaMemcpyAsync(dz + zb + rb, s, rb, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (kb) HK_CUDA(cudaMemcpyAsync(dz + zb + 2 * rb, kappas, kb, cudaMemcpyHostToDevice, ctx->copy_stream));
            HK_CUDA(cudaEventRecord(ctx->ev_h2d[slot], ctx->copy_stream));
            HK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[slot], 0));
            const size_t out_bytes = K * (2 * p1 + p2 + 3);
            char* d_out = (char*)ctx->ws_io[3].ensure(out_bytes + 64);
            char* dA = d_out;
            char* dB = dA + K * p1;
            char* dC = dB + K * p2;
            uint8_t* d_inf = (uint8_t*)(dC + K * p1);
            const uint64_t* dzz = (const uint64_t*)dz;
            const uint64_t* dr = (const uint64_t*)(dz + zb);
            const uint64_t* ds = (const uint64_t*)(dz + zb + rb);
            const uint64_t* dk = (const uint64_t*)(dz + zb + 2 * rb);
            by_curve(ctx, [&] { prove_run<Bls12381>(ctx, pk, K, dzz, dr, ds, dk, dA, dB, dC, d_inf); },
                     [&] { prove_run<Bn254>(ctx, pk, K, dzz, dr, ds, dk, dA, dB, dC, d_inf); });
            HK_CUDA(cudaEventRecord(ctx->ev_in_free[slot], ctx->stream));
            HK_CUDA(cudaMemcpyAsync(a_affine, dA, K * p1, cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(cudaMemcpyAsync(b_affine, dB, K * p2, cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(cudaMemcpyAsync(c_affine, dC, K * p1, cudaMemcpyDeviceToHost, ctx->stream));
            HK_CUDA(cudaMemcpyAsync(inf, d_inf, K * 3, cudaMemcpyDeviceToHost, ctx->stream));
            return;
        }
        // host inputs -> device [z | r | s | kappas]; pageable inputs go through a pinned staging buffer
        // one proof at a time so the DMA of proof k overlaps the staging copy of proof k+1; pinned
        // inputs are DMA'd in place.  One D2H copy brings all proofs back.
        HK_REQUIRE(memspace == HK_HOST || memspace == HK_HOST_PINNED, HK_ERR_INVALID, "bad memspace");
        const size_t zb1 = pk.m * 32, zb = K * zb1, rb = K * 32, kb = K * (size_t)nk * 32;
        const size_t total = zb + 2 * rb + kb;
        const size_t out_bytes = K * (2 * p1 + p2 + 3);
        char* dz = (char*)ctx->ws_z.ensure(total + 64);
        char* pin;
        char* pout;
        if (memspace == HK_HOST_PINNED) {
            pin = nullptr;
            pout = (char*)ctx->pinned_buf(out_bytes + 64);
            HK_CUDA(cudaMemcpyAsync(dz, z, zb, cudaMemcpyHostToDevice, ctx->stream));
            HK_CUDA(cudaMemcpyAsync(dz + zb, r, rb, cudaMemcpyHostToDevice, ctx->stream));
            HK_CUDA(cudaMemcpyAsync(dz + zb + rb, s, rb, cudaMemcpyHostToDevice, ctx->stream));
            if (kb) HK_CUDA(cudaMemcpyAsync(dz + zb + 2 * rb, kappas, kb, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            pin = (char*)ctx->pinned_buf(total + out_bytes + 64);
            pout = pin + total;
            memcpy(pin + zb, r, rb);
            memcpy(pin + zb + rb, s, rb);
            if (kb) memcpy(pin + zb + 2 * rb, kappas, kb);
            HK_CUDA(cudaMemcpyAsync(dz + zb, pin + zb, 2 * rb + kb, cudaMemcpyHostToDevice, ctx->stream));
            for (size_t k = 0; k < K; k++) {
                memcpy(pin + k * zb1, (const char*)z + k * zb1, zb1);
                HK_CUDA(cudaMemcpyAsync(dz + k * zb1, pin + k * zb1, zb1, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        char* d_out = (char*)ctx->ws_io[3].ensure(out_bytes + 64);
        char* dA = d_out;
        char* dB = dA + K * p1;
        char* dC = dB + K * p2;
        uint8_t* d_inf = (uint8_t*)(dC + K * p1);
        const uint64_t* dzz = (const uint64_t*)dz;
        const uint64_t* dr = (const uint64_t*)(dz + zb);
        const uint64_t* ds = (const uint64_t*)(dz + zb + rb);
        const uint64_t* dk = (const uint64_t*)(dz + zb + 2 * rb);
        by_curve(ctx, [&] { prove_run<Bls12381>(ctx, pk, K, dzz, dr, ds, dk, dA, dB, dC, d_inf); },
                 [&] { prove_run<Bn254>(ctx, pk, K, dzz, dr, ds, dk, dA, dB, dC, d_inf); });
        HK_CUDA(cudaMemcpyAsync(pout, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(a_affine, pout, K * p1);
        memcpy(b_affine, pout + K * p1, K * p2);
        memcpy(c_affine, pout + K * (p1 + p2), K * p1);
        if (inf) memcpy(inf, pout + K * (2 * p1 + p2), K * 3);
    });
}

HK_EXPORT int32_t hk_prove(hk_ctx* ctx, uint64_t pk_h, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                           const uint64_t* kappas, uint32_t memspace, void* a_affine, void* b_affine, void* c_affine,
                           uint8_t* inf) {
    uint8_t scratch[3];
    // the flags are written asynchronously for HK_DEVICE / HK_HOST_PINNED_ASYNC: a stack scratch would dangle
    if ((memspace == HK_DEVICE || memspace == HK_HOST_PINNED_ASYNC) && !inf) return HK_ERR_INVALID;
    return hk_prove_batch(ctx, pk_h, 1, z, r, s, kappas, memspace, a_affine, b_affine, c_affine,
                          inf ? inf : (memspace == HK_DEVICE ? nullptr : scratch));
}

HK_EXPORT int32_t hk_field_op(hk_ctx* ctx, uint32_t field, uint32_t op, const uint64_t* a, const uint64_t* b,
                              uint64_t* out, size_t n) {
    return guarded(ctx, [&] {
        HK_REQUIRE(a && out, HK_ERR_INVALID, "NULL argument");
        HK_REQUIRE(op <= HK_FOP_MULSUB2, HK_ERR_INVALID, "bad op");
        HK_REQUIRE(b || (op >= HK_FOP_INV && op <= HK_FOP_SQR), HK_ERR_INVALID, "b is NULL");
        if (n == 0) return;
        bool bls = ctx->curve == HK_CURVE_BLS12_381;
        if (field == HK_FIELD_FR) {
            if (bls) field_op_run<Bls12381::Fr>(ctx, op, a, b, out, n);
            else field_op_run<Bn254::Fr>(ctx, op, a, b, out, n);
        } else if (field == HK_FIELD_FQ) {
            if (bls) field_op_run<Bls12381::Fq>(ctx, op, a, b, out, n);
            else field_op_run<Bn254::Fq>(ctx, op, a, b, out, n);
        } else if (field == HK_FIELD_FQ2) {
            if (bls) field_op_run<Bls12381::Fq2>(ctx, op, a, b, out, n);
            else field_op_run<Bn254::Fq2>(ctx, op, a, b, out, n);
        } else {
            throw HkError{HK_ERR_INVALID, "bad field id"};
        }
    });
}

HK_EXPORT int32_t hk_profile_enable(hk_ctx* ctx, int32_t on) {
    return guarded(ctx, [&] {
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->prof_on = on != 0;
        ctx->serial_lanes = on == 2;
    });
}
HK_EXPORT int32_t hk_profile_reset(hk_ctx* ctx) {
    return guarded(ctx, [&] {
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
        prof_clear(ctx);
    });
}
HK_EXPORT int32_t hk_profile_read(hk_ctx* ctx, uint32_t kernel_id, double* total_ms, uint64_t* launches,
                                  uint64_t* units) {
    return guarded(ctx, [&] {
        HK_REQUIRE(kernel_id < HK_PROF_COUNT && total_ms && launches && units, HK_ERR_INVALID, "bad argument");
        HK_CUDA(cudaStreamSynchronize(ctx->stream));
        double ms = 0;
        uint64_t n = 0, u = 0;
        for (auto& r : ctx->prof) {
            if (r.id != (int)kernel_id) continue;
            float t = 0;
            HK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
            ms += t;
            n++;
            u += r.pinned_units ? (uint64_t)*r.pinned_units : r.static_units;
        }
        *total_ms = ms;
        *launches = n;
        *units = u;
    });
}

HK_EXPORT int32_t hk_bench_field_mul(hk_ctx* ctx, uint32_t field, uint32_t threads, uint32_t iters, float* ms) {
    return guarded(ctx, [&] {
        HK_REQUIRE(ms && threads && iters, HK_ERR_INVALID, "bad argument");
        bool bls = ctx->curve == HK_CURVE_BLS12_381;
        threads = (threads + 255) / 256 * 256;
        void* out = ctx->ws_io[0].ensure((size_t)threads * 48);
        auto run = [&](auto tag) {
            using F = decltype(tag);
            k_bench_mul<F><<<threads / 256, 256, 0, ctx->stream>>>((F*)out, 2);   // warm-up
            *ms = time_ms(ctx, [&] {
                k_bench_mul<F><<<threads / 256, 256, 0, ctx->stream>>>((F*)out, iters);
                ctx->launches++;
            });
            HK_CUDA(cudaGetLastError());
        };
        if (field == HK_FIELD_FR) {
            if (bls) run(Bls12381::Fr{});
            else run(Bn254::Fr{});
        } else if (field == HK_FIELD_FQ) {
            if (bls) run(Bls12381::Fq{});
            else run(Bn254::Fq{});
        } else {
            throw HkError{HK_ERR_INVALID, "bench: field must be FR or FQ"};
        }
    });
}

HK_EXPORT int32_t hk_bench_imad_wide(hk_ctx* ctx, uint32_t iters, double* macs_per_s) {
    return guarded(ctx, [&] {
        HK_REQUIRE(macs_per_s && iters, HK_ERR_INVALID, "bad argument");
        const uint32_t blocks = ctx->sm_count * 8, tpb = 256;
        uint64_t* out = ctx->ws_io[0].as<uint64_t>((size_t)blocks * tpb);
        k_bench_imad_wide<<<blocks, tpb, 0, ctx->stream>>>(out, 4);
        float ms = time_ms(ctx, [&] {
            k_bench_imad_wide<<<blocks, tpb, 0, ctx->stream>>>(out, iters);
            ctx->launches++;
        });
        HK_CUDA(cudaGetLastError());
        double macs = (double)blocks * tpb * (double)iters * 32.0;
        *macs_per_s = macs / (ms * 1e-3);
    });
}
