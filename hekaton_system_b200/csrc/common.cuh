// Context, error handling, workspace management shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "../../include/hekaton_b200.h"
#include "ec.cuh"

namespace hk {

// ---------------------------------------------------------------------------------------
// curve configurations
// ---------------------------------------------------------------------------------------
struct Bls12381 {
    using Fq = Fp<bls12_381_fq_tag>;
    using Fr = Fp<bls12_381_fr_tag>;
    using Fq2 = Fp2T<Fq>;
    static constexpr uint32_t ID = HK_CURVE_BLS12_381;
    static constexpr int FR_BITS = 255;
    static constexpr int TWO_ADICITY = 32;
    static constexpr uint64_t FR_GENERATOR = 7;
};
struct Bn254 {
    using Fq = Fp<bn254_fq_tag>;
    using Fr = Fp<bn254_fr_tag>;
    using Fq2 = Fp2T<Fq>;
    static constexpr uint32_t ID = HK_CURVE_BN254;
    static constexpr int FR_BITS = 254;
    static constexpr int TWO_ADICITY = 28;
    static constexpr uint64_t FR_GENERATOR = 5;
};

struct Scalar256 {
    uint64_t v[4];
};

// coordinate field of G1 / G2
template <class Curve, int GROUP>
struct GroupFieldT;
template <class Curve>
struct GroupFieldT<Curve, 1> {
    using F = typename Curve::Fq;
};
template <class Curve>
struct GroupFieldT<Curve, 2> {
    using F = typename Curve::Fq2;
};

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
struct HkError {
    int32_t code;
    std::string msg;
};

#define HK_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            throw hk::HkError{HK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                               " (" __FILE__ ":" + std::to_string(__LINE__) + ")"}; \
        }                                                                                      \
    } while (0)

#define HK_REQUIRE(cond, code, text)                    \
    do {                                                \
        if (!(cond)) throw hk::HkError{(code), (text)}; \
    } while (0)

// ---------------------------------------------------------------------------------------
// grow-only device buffer
// ---------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void* ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) {
                cudaFree(p);
                p = nullptr;
                cap = 0;
            }
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                p = nullptr;
                throw HkError{HK_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(want) + "): " +
                                                cudaGetErrorString(e)};
            }
            cap = want;
        }
        return p;
    }
    template <class T>
    T* as(size_t count) {
        return reinterpret_cast<T*>(ensure(count * sizeof(T)));
    }
};

struct BaseSet {
    int group = 1;          // 1 = G1, 2 = G2
    size_t n = 0;
    size_t point_bytes = 0;
    void* d = nullptr;      // affine points, AoS
    // optional window-shifted table (hk_bases_precompute / hk_pk_create):
    //   tbl[w * n + i] = 2^(tbl_c * w) * P_i   (affine), w < tbl_W.  Trades HBM capacity for the
    //   removal of the serial window-combination tail and a W-fold smaller bucket set.
    void* tbl = nullptr;
    int tbl_c = 0, tbl_W = 0;
    uint8_t* skip = nullptr;   // 1 byte per point: base is the point at infinity (built with the table)
    std::mutex tbl_mu;         // serialises table construction (forked contexts share base sets); tbl is published last
    BaseSet() = default;
    BaseSet(const BaseSet&) = delete;
    BaseSet& operator=(const BaseSet&) = delete;
    ~BaseSet() {
        if (d) cudaFree(d);
        if (tbl) cudaFree(tbl);
        if (skip) cudaFree(skip);
    }
};

// What the MSM pipeline needs to know about a base set.
struct MsmBases {
    const void* pts = nullptr;   // affine points
    size_t n_total = 0;          // points in the set (table row stride)
    const void* tbl = nullptr;   // window-shifted table or nullptr
    int c = 0, W = 0;            // table geometry
    const uint8_t* skip = nullptr;   // per-point "is infinity" flags or nullptr
};
inline MsmBases msm_view(const BaseSet& b) { return MsmBases{b.d, b.n, b.tbl, b.tbl_c, b.tbl_W, b.skip}; }

// Window geometry of one MSM.
struct MsmPlan {
    int c;       // window bits
    int W;       // windows
    uint32_t B;  // buckets per window = 2^(c-1)
};

// A key-sorted digit list produced by msm_sort (msm_impl.cuh); valid until the owning lane sorts again.
struct MsmSorted {
    MsmPlan pl{};
    bool table = false;
    size_t n = 0, K = 0, offset = 0, n_total = 0;
    uint32_t Wred = 0, nsets = 0, item_stride = 0;
    size_t nkeys = 0, max_items = 0;
    const uint64_t* pairs = nullptr;     // key << 32 | table index << 1 | sign, key order
    const uint32_t* d_count = nullptr;   // device: number of pairs
    int lane = 0;                        // lane whose workspace holds the list (ev_sorted[lane] marks it ready)
};

struct ProfRec {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int id = 0;
    uint64_t static_units = 0;
    uint32_t* pinned_units = nullptr;   // filled by an async D2H copy of a device counter
};

// Scratch of one MSM pipeline; a context has one per lane (main stream, side stream) so two MSMs
// of the same proof batch can run concurrently.
struct MsmWs {
    DevBuf scalars, hist, items, buckets, chunks, wsum, scan, misc;
};

struct NttDomain;   // ntt.cu
struct R1csDev;     // qap.cu
struct PkDev;       // prove.cu

// Resources shared by a context and its forks (hk_ctx_fork): everything that is uploaded once
// and then only read by kernels.  Host-side access is serialised by `mu`.
struct SharedState {
    std::mutex mu;
    uint64_t next_handle = 1;
    std::map<uint64_t, std::shared_ptr<BaseSet>> bases;
    std::map<uint32_t, std::shared_ptr<NttDomain>> domains;   // by log_n
    std::map<uint64_t, std::shared_ptr<R1csDev>> r1cs;
    std::map<uint64_t, std::shared_ptr<PkDev>> pks;
};

}  // namespace hk

struct hk_ctx {
    int device = 0;
    uint32_t curve = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // Lanes of a proof batch (prove.cu).  Lane 0 is the caller's stream.  Every other lane has a
    // CONTROL stream of the highest priority for its short latency-bound kernels (sort, folds,
    // reduction, NTT passes) and a BULK stream for its bucket-accumulation kernel, whose priority falls
    // with the lane number: each accumulation fills the GPU on its own, so the bulk priorities fix the
    // order in which the block scheduler serves them, while the control kernels of all lanes find
    // slots as blocks retire and hide under whichever accumulation is running.
    static constexpr int LANES = 6;
    cudaStream_t lane_ctl[LANES] = {}, lane_bulk[LANES] = {};
    cudaEvent_t ev_fork = nullptr, ev_sorted[LANES] = {}, ev_ready[LANES] = {}, ev_accum[LANES] = {}, ev_lane_done[LANES] = {};
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    std::shared_ptr<hk::SharedState> sh;

    // scratch (grow-only; a context is single-threaded by contract)
    hk::MsmWs mws[LANES];              // one MSM scratch per lane
    hk::DevBuf ws_out;
    hk::DevBuf ws_ntt_tmp, ws_qap[4], ws_io[6], ws_z, ws_misc;
    // serial_lanes (hk_profile_enable(ctx, 2)): every lane maps to the caller's stream, so kernels run one
    // at a time and the profiling events bracket each kernel alone (roofline measurements)
    bool serial_lanes = false;
    // MSM tuning knobs (hk_ctx_set_option; 0 = built-in default, see msm_accumulate)
    uint32_t opt_accum_waves = 0, opt_accum_smin = 0, opt_fold_seg = 0, opt_reduce_g = 0, opt_sort_global = 0, opt_g2_pair = 0, opt_quad_max = 0, opt_no_fold_resolve = 0;
    cudaStream_t lane_stream(int lane) const { return !serial_lanes && lane > 0 && lane < LANES ? lane_ctl[lane] : stream; }
    cudaStream_t bulk_stream(int lane) const { return !serial_lanes && lane > 0 && lane < LANES ? lane_bulk[lane] : stream; }
    // stream used by the NTT / QAP kernels (normally the main stream; prove_run points it at the H lane)
    cudaStream_t qap_stream = nullptr;
    cudaStream_t qstream() const { return qap_stream ? qap_stream : stream; }
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    // asynchronous host-buffer calls (HK_HOST_PINNED_ASYNC): inputs are DMA'd on a copy stream into one of two
    // device staging buffers, so the upload of call j+1 overlaps the kernels of call j
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {}, ev_in_free[2] = {};
    uint64_t async_seq = 0;
    hk::DevBuf ws_zin[2];

    // roofline profiling (hk_profile_*)
    bool prof_on = false;
    std::vector<hk::ProfRec> prof;
    uint32_t* prof_pinned = nullptr;
    size_t prof_pinned_used = 0;
    static constexpr size_t PROF_SLOTS = 1 << 16;

    void* pinned_buf(size_t bytes) {
        if (bytes > pinned_cap) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr;
            pinned_cap = 0;
            if (cudaMallocHost(&pinned, bytes + 256) != cudaSuccess)
                throw hk::HkError{HK_ERR_NOMEM, "cudaMallocHost failed"};
            pinned_cap = bytes + 256;
        }
        return pinned;
    }
};

namespace hk {

inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// NVTX range over the ENQUEUE of a pipeline stage, named after the reference's timers (ark_std start_timer! labels in
// cp-groth16/src/prover.rs:58-150 and committer.rs:59-96), so an nsys / ncu timeline of this library lines up with
// the reference's --features print-trace output.  The kernels run asynchronously on the lane streams; tools
// attribute them to the range that launched them.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Wraps every extern "C" body: activates the device, maps exceptions to codes.
template <class Fn>
inline int32_t guarded(hk_ctx* ctx, Fn&& fn) {
    if (!ctx) return HK_ERR_INVALID;
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) throw HkError{HK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
        fn();
        return HK_OK;
    } catch (const HkError& e) {
        ctx->err = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return HK_ERR_INVALID;
    } catch (...) {
        ctx->err = "unknown exception";
        return HK_ERR_INVALID;
    }
}

#define HK_LAUNCH_CHECK(ctx)       \
    do {                           \
        (ctx)->launches++;         \
        HK_CUDA(cudaGetLastError()); \
    } while (0)

// Bracket one kernel launch with events when profiling is on.
inline int prof_begin(hk_ctx* ctx, int id, cudaStream_t st = nullptr) {
    if (!ctx->prof_on) return -1;
    if (!st) st = ctx->stream;
    ProfRec r;
    r.id = id;
    HK_CUDA(cudaEventCreate(&r.e0));
    HK_CUDA(cudaEventCreate(&r.e1));
    HK_CUDA(cudaEventRecord(r.e0, st));
    ctx->prof.push_back(r);
    return (int)ctx->prof.size() - 1;
}
inline void prof_end(hk_ctx* ctx, int idx, uint64_t static_units, const uint32_t* d_units, cudaStream_t st = nullptr) {
    if (idx < 0) return;
    if (!st) st = ctx->stream;
    ProfRec& r = ctx->prof[idx];
    HK_CUDA(cudaEventRecord(r.e1, st));
    r.static_units = static_units;
    if (d_units) {
        if (!ctx->prof_pinned) HK_CUDA(cudaMallocHost((void**)&ctx->prof_pinned, hk_ctx::PROF_SLOTS * sizeof(uint32_t)));
        if (ctx->prof_pinned_used < hk_ctx::PROF_SLOTS) {
            r.pinned_units = ctx->prof_pinned + ctx->prof_pinned_used++;
            HK_CUDA(cudaMemcpyAsync(r.pinned_units, d_units, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
    }
}

// ---- per-module entry points (implemented in the respective .cu files) -----------------
// msm_sort / msm_accumulate: the two halves of msm_run, exposed so that one sorted list can feed several base
// sets of the same geometry (prove.cu: B in G1 and in G2).
template <class Curve>
MsmSorted msm_sort(hk_ctx* ctx, const MsmBases& bases, size_t offset, size_t n, const Scalar256* d_scalars, size_t K,
                   bool scalars_mont, size_t sc_stride, int lane);
template <class Curve, int GROUP>
void msm_accumulate(hk_ctx* ctx, const MsmSorted& sorted, const MsmBases& bases, void* d_out_jac, int lane);
// msm: out[k] = sum_{i<n} scalars[k * sc_stride + i] * bases[offset + i] for k < K, in ONE pass (the
// batch index is an extra bucket-set dimension).  Scalars are device pointers; out receives K
// Jacobian points (device).  sc_stride = 0 means n.
template <class Curve, int GROUP>
void msm_run(hk_ctx* ctx, const MsmBases& bases, size_t offset, size_t n, const Scalar256* d_scalars, size_t K,
             bool scalars_mont, void* d_out_jac, size_t sc_stride = 0, int lane = 0);
// build the window-shifted table of a base set (no-op if present); density = expected fraction of
// full-width scalars (1.0 = uniform), steers the window size
template <class Curve, int GROUP>
void bases_precompute(hk_ctx* ctx, BaseSet& bs, double density);

template <class Curve>
void ntt_run(hk_ctx* ctx, uint64_t* d_data, uint32_t log_n, size_t K, uint32_t dir, uint32_t domain);

}  // namespace hk
