// Pippenger MSM for sm_100a.  Replaces ark-ec 0.4 `VariableBaseMSM::msm_bigint` as called from
// /root/reference/cp-groth16/src/prover.rs:117,129,168 and committer.rs:89,113.
//
// Pipeline (all on the lane's streams, no host synchronisation):
//   msm_sort        1. k_msm_hist      signed-digit recoding of every scalar (window c, digits in
//                                      [-2^(c-1), 2^(c-1)], top window unrecoded) + histogram of
//                                      (window, |digit|) keys
//                   2. exclusive scan  bucket start offsets
//                   3. k_msm_scatter   counting sort: pair = key << 32 | point index << 1 | sign
//   msm_accumulate  4. k_msm_accum     balanced segmented accumulation: every thread owns one segment of the
//                                      key-sorted list (so skewed 0/1-heavy witness scalars cost the same as
//                                      uniform ones).  Interior runs go straight to their bucket, the two runs cut
//                                      by the segment borders go to a partial list.  Default: 32 items per thread,
//                                      as many CTAs as that takes (fine-grained CTAs balance the tail best);
//                                      HK_OPT_ACCUM_WAVES caps the grid, then the segment length follows from the
//                                      item count ON THE DEVICE (S = max(S_min, ceil(count / threads))).
//                   5. k_msm_fold      the same scheme over the partial list (full XYZZ additions), level by level
//                   6. k_msm_reduce_level / k_msm_tree_sum   sum_v v * B_v by a RECURSIVE chunked weighted sum:
//                                      chunk t of G buckets gives acc_t = sum_j j B_(tG+j) and run_t = sum_j B_(tG+j);
//                                      sum_v v B_v = sum_t acc_t + G * sum_t t run_t, and the second sum is the same
//                                      problem on T = B/G points.  2 B (1 + 1/G + ...) additions and no scalar
//                                      multiplication; the per-level sums A_l = sum_t acc_t come from one tree sum
//                                      and are combined as A_0 + G (A_1 + G (A_2 + ...)).
//                   7. k_msm_combine   Horner over the levels and (without a table) over the windows -> Jacobian
// The sorted list of one call can feed several accumulations over DIFFERENT base sets of the same geometry
// (prove.cu: B in G1 and B in G2 use the same scalars -- one sort instead of two).
//
// Table mode (base sets with a window-shifted table, tbl[w*n+i] = 2^(c w) P_i): every window's
// digit selects from its own table row, so ALL windows share ONE bucket set: keys are bucket
// values only, the reduce stage is W times smaller and step 7 has no window doublings at all.
// Memory cost: W x the base set (HBM is 180 GB).
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "fold_runs.cuh"

namespace hk {

inline MsmPlan msm_plan_for(int c, int scalar_bits) {
    return {c, (scalar_bits + 1 + c - 1) / c, 1u << (c - 1)};
}

// Window size from a cost model in Fq multiplications, calibrated on B200 (profiles/r01b, r02):
//   accumulate: 10 per item; items = density * n per window (+ ~n/2 once for 0/1 scalars when
//               density < 1: a scalar equal to 1 has a single non-zero digit)
//   reduce:     ~300 per bucket (these kernels are latency-bound, far from the multiplier peak), per window
//               without a table, once with a table
// `density` = expected fraction of scalars that are "full width" (witness vectors of SHA-256 /
// Poseidon circuits are mostly bits: SURVEY.md 8d).
inline MsmPlan msm_plan(size_t n, int scalar_bits, bool table, double density = 1.0) {
    double best = 1e300;
    MsmPlan p{4, 0, 0};
    if (density <= 0 || density > 1) density = 1.0;
    const double ones = density < 1.0 ? 0.5 * double(n) : 0.0;
    for (int c = 3; c <= (table ? 20 : 16); c++) {
        MsmPlan q = msm_plan_for(c, scalar_bits);
        double B = double(q.B);
        double acc = 10.0 * (density * double(n) * q.W + ones);
        double cost = table ? acc + B * 300.0 : acc + q.W * B * 300.0;
        if (cost < best) {
            best = cost;
            p = q;
        }
    }
    return p;
}

// signed digit of window w; `carry` threads through windows 0..W-1 in order
__device__ __forceinline__ int msm_digit(const uint64_t s[4], int w, int c, int W, uint32_t& carry) {
    int bit = w * c;
    uint64_t raw = 0;
    if (bit < 256) {
        int limb = bit >> 6, sh = bit & 63;
        raw = s[limb] >> sh;
        if (sh + c > 64 && limb + 1 < 4) raw |= s[limb + 1] << (64 - sh);
    }
    uint32_t coef = (uint32_t)(raw & ((1u << c) - 1)) + carry;
    if (w == W - 1) {
        // canonical scalars (< 2^scalar_bits) give coef <= 2^(c-1) here; anything larger is a caller error and is
        // clamped so that it can never index past the bucket array
        carry = 0;
        const uint32_t top = 1u << (c - 1);
        return (int)(coef > top ? top : coef);
    }
    carry = (coef + (1u << (c - 1))) >> c;
    return (int)coef - (int)(carry << c);
}

template <class Fr>
__global__ void k_scalars_from_mont(const Scalar256* in, Scalar256* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x;
#pragma unroll
    for (int k = 0; k < 4; k++) x.r.v[k] = in[i].v[k];
    x = x.from_mont();
#pragma unroll
    for (int k = 0; k < 4; k++) out[i].v[k] = x.r.v[k];
}

// Bucket sets: one per (vector k, window w) without a table, one per vector k with a table.
//   key  = (k * sets_per_vec + w * per_window) * B + (|d| - 1)      per_window = 1 / 0
//   item = w * item_stride + i                                      item_stride = 0 / n_total
// The K scalar vectors of a batch share the bases, so the batch index is just another bucket-set
// dimension; `sc_stride` is the distance (in scalars) between consecutive vectors.
// Both sort kernels aggregate their atomics per warp: lanes whose current digit falls into the same
// bucket are found with match.any, one lane issues a single atomic for the group.  Witness vectors of
// SHA-256 / Poseidon circuits are mostly bits (45 % of all scalars are 1), so without this almost half of
// the atomics of a launch hit the same few addresses and serialise in L2.
static __global__ void k_msm_hist(const Scalar256* sc, size_t n, size_t K, size_t sc_stride, int c, int W, uint32_t B,
                                  uint32_t sets_per_vec, uint32_t per_window, const uint8_t* __restrict__ skip,
                                  uint32_t* hist) {
    const unsigned lane = threadIdx.x & 31u;
    const size_t total = n * K, padded = (total + 31) & ~(size_t)31;      // whole warps take part in match.any
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < padded; g += (size_t)gridDim.x * blockDim.x) {
        bool active = g < total;
        size_t k = 0, i = 0;
        uint64_t s[4] = {0, 0, 0, 0};
        if (active) {
            k = g / n;
            i = g - k * n;
            if (skip && skip[i]) active = false;          // base at infinity contributes nothing
        }
        if (active) {
            const Scalar256* p = sc + k * sc_stride + i;
            s[0] = p->v[0], s[1] = p->v[1], s[2] = p->v[2], s[3] = p->v[3];
            active = (s[0] | s[1] | s[2] | s[3]) != 0;
        }
        if (!__any_sync(0xffffffffu, active)) continue;
        uint32_t carry = 0;
        const size_t base = k * sets_per_vec;
        for (int w = 0; w < W; w++) {
            int d = active ? msm_digit(s, w, c, W, carry) : 0;
            const unsigned amask = __ballot_sync(0xffffffffu, d != 0);
            if (d != 0) {
                uint32_t key = (uint32_t)((base + (size_t)w * per_window) * B) + (uint32_t)(abs(d) - 1);
                unsigned peers = __match_any_sync(amask, key);
                if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[key], (uint32_t)__popc(peers));
            }
        }
    }
}

// pairs[pos] = key << 32 | item: the accumulation kernel reads one 8-byte stream instead of two 4-byte ones
static __global__ void k_msm_scatter(const Scalar256* sc, size_t n, size_t K, size_t sc_stride, int c, int W, uint32_t B,
                                     uint32_t sets_per_vec, uint32_t per_window, uint32_t item_stride,
                                     const uint8_t* __restrict__ skip, uint32_t* cursor, uint64_t* pairs) {
    const unsigned lane = threadIdx.x & 31u;
    const size_t total = n * K, padded = (total + 31) & ~(size_t)31;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < padded; g += (size_t)gridDim.x * blockDim.x) {
        bool active = g < total;
        size_t k = 0, i = 0;
        uint64_t s[4] = {0, 0, 0, 0};
        if (active) {
            k = g / n;
            i = g - k * n;
            if (skip && skip[i]) active = false;
        }
        if (active) {
            const Scalar256* p = sc + k * sc_stride + i;
            s[0] = p->v[0], s[1] = p->v[1], s[2] = p->v[2], s[3] = p->v[3];
            active = (s[0] | s[1] | s[2] | s[3]) != 0;
        }
        if (!__any_sync(0xffffffffu, active)) continue;
        uint32_t carry = 0;
        const size_t base = k * sets_per_vec;
        for (int w = 0; w < W; w++) {
            int d = active ? msm_digit(s, w, c, W, carry) : 0;
            const unsigned amask = __ballot_sync(0xffffffffu, d != 0);
            if (d != 0) {
                uint32_t key = (uint32_t)((base + (size_t)w * per_window) * B) + (uint32_t)(abs(d) - 1);
                unsigned peers = __match_any_sync(amask, key);
                const int leader = __ffs(peers) - 1;
                uint32_t pos0 = 0;
                if ((int)lane == leader) pos0 = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
                pos0 = __shfl_sync(peers, pos0, leader);
                uint32_t pos = pos0 + (uint32_t)__popc(peers & ((1u << lane) - 1u));
                uint32_t item = (((uint32_t)w * item_stride + (uint32_t)i) << 1) | (d < 0 ? 1u : 0u);
                pairs[pos] = ((uint64_t)key << 32) | item;
            }
        }
    }
}

// ---- counting sort with block-private histograms in shared memory ---------------------------------
// With a window table all windows of a scalar vector share ONE bucket set, so a vector's whole key space is
// B = 2^(c-1) counters: 16 KB for the assignment MSMs (c = 13), 128 KB for the H MSM (c = 16) -- it fits in the
// 227 KB of shared memory of an SM.  Block (j, k) owns chunk j of vector k: it histograms its digits with
// shared-memory atomics, publishes cnt[(k B + b) J + j], and after ONE exclusive scan over that array (key-major,
// block-minor: the offsets of a key's items from block j follow those from block j - 1) it re-reads its chunk and
// places every item with a shared-memory cursor.  No global atomic is issued at all; the global-atomic kernels above
// remain for bucket sets that do not fit (no table: W B counters).
// Window 0 aggregates equal digits of a warp first (match.any): witness vectors are mostly ones, whose single
// non-zero digit would otherwise serialise a warp on one shared-memory word.
template <bool SCATTER>
__global__ void __launch_bounds__(1024)
k_msm_sort_smem(const Scalar256* __restrict__ sc, size_t n, size_t sc_stride, int c, int W, uint32_t B, uint32_t J,
                uint32_t chunk, uint32_t item_stride, const uint8_t* __restrict__ skip, uint32_t* cnt, uint64_t* pairs) {
    extern __shared__ uint32_t sm_cnt[];
    const uint32_t j = blockIdx.x, k = blockIdx.y;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t* mine = cnt + (size_t)k * B * J + j;              // stride J between consecutive buckets
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) sm_cnt[b] = SCATTER ? mine[(size_t)b * J] : 0u;
    __syncthreads();
    const size_t lo = (size_t)j * chunk;
    const size_t hi = lo + chunk < n ? lo + chunk : n;
    const size_t span = hi > lo ? hi - lo : 0;
    const size_t padded = (span + 31) & ~(size_t)31;           // whole warps take part in match.any
    const uint32_t key_base = k * B;
    for (size_t o = threadIdx.x; o < padded; o += blockDim.x) {
        const size_t i = lo + o;
        bool active = o < span;
        uint64_t s[4] = {0, 0, 0, 0};
        if (active && skip && skip[i]) active = false;          // base at infinity contributes nothing
        if (active) {
            const Scalar256* p = sc + (size_t)k * sc_stride + i;
            s[0] = p->v[0], s[1] = p->v[1], s[2] = p->v[2], s[3] = p->v[3];
            active = (s[0] | s[1] | s[2] | s[3]) != 0;
        }
        if (!__any_sync(0xffffffffu, active)) continue;
        uint32_t carry = 0;
        for (int w = 0; w < W; w++) {
            const int d = active ? msm_digit(s, w, c, W, carry) : 0;
            uint32_t pos = 0;
            if (w == 0) {
                const unsigned amask = __ballot_sync(0xffffffffu, d != 0);
                if (d != 0) {
                    const uint32_t b = (uint32_t)(abs(d) - 1);
                    const unsigned peers = __match_any_sync(amask, b);
                    const int leader = __ffs(peers) - 1;
                    uint32_t pos0 = 0;
                    if ((int)lane == leader) pos0 = atomicAdd(&sm_cnt[b], (uint32_t)__popc(peers));
                    if (SCATTER) pos = __shfl_sync(peers, pos0, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
                }
            } else if (d != 0) {
                pos = atomicAdd(&sm_cnt[abs(d) - 1], 1u);
            }
            if (SCATTER && d != 0) {
                const uint32_t item = (((uint32_t)w * item_stride + (uint32_t)i) << 1) | (d < 0 ? 1u : 0u);
                pairs[pos] = ((uint64_t)(key_base + (uint32_t)(abs(d) - 1)) << 32) | item;
            }
        }
    }
    if (!SCATTER) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) mine[(size_t)b * J] = sm_cnt[b];
    }
}

// ---- exclusive scan (u32), in place; total written to *total --------------------------------
static constexpr int SCAN_TPB = 256;
static constexpr int SCAN_IPT = 8;
static constexpr int SCAN_TILE = SCAN_TPB * SCAN_IPT;

static __global__ void k_scan_tiles(uint32_t* data, size_t len, uint32_t* tile_sums) {
    __shared__ uint32_t sh[SCAN_TPB];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) {
        v[k] = (base + k < len) ? data[base + k] : 0;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the per-thread sums
    for (int off = 1; off < SCAN_TPB; off <<= 1) {
        uint32_t t = (threadIdx.x >= off) ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t excl = sh[threadIdx.x] - s;
    if (threadIdx.x == SCAN_TPB - 1) tile_sums[blockIdx.x] = sh[threadIdx.x];
    uint32_t run = excl;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) {
        if (base + k < len) data[base + k] = run;
        run += v[k];
    }
}

static __global__ void k_scan_sums(uint32_t* tile_sums, size_t ntiles, uint32_t* total) {
    // single block, sequential over chunks of blockDim
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < ntiles; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        uint32_t v = (i < ntiles) ? tile_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < (int)blockDim.x; off <<= 1) {
            uint32_t t = (threadIdx.x >= (unsigned)off) ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t incl = sh[threadIdx.x];
        if (i < ntiles) tile_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

static __global__ void k_scan_add(uint32_t* data, size_t len, const uint32_t* tile_sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++)
        if (base + k < len) data[base + k] += add;
}

inline void exclusive_scan_u32(hk_ctx* ctx, MsmWs& ws, cudaStream_t st, uint32_t* d, size_t len, uint32_t* d_total) {
    size_t ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* sums = ws.scan.as<uint32_t>(ntiles + 1);
    k_scan_tiles<<<(unsigned)ntiles, SCAN_TPB, 0, st>>>(d, len, sums);
    HK_LAUNCH_CHECK(ctx);
    k_scan_sums<<<1, 1024, 0, st>>>(sums, ntiles, d_total);
    HK_LAUNCH_CHECK(ctx);
    k_scan_add<<<(unsigned)ntiles, SCAN_TPB, 0, st>>>(d, len, sums);
    HK_LAUNCH_CHECK(ctx);
}

// ---- bucket accumulation ---------------------------------------------------------------------
// Level 0: thread t owns items [t*S, (t+1)*S) of the key-sorted list (mixed adds of affine bases), with
// S = max(s_min, ceil(count / nthreads)) computed from the device-side item count.  Runs strictly inside the
// segment are complete -> bucket array.  The first and the last run of a segment may continue in a
// neighbour -> partial list slots 2t, 2t+1 (key = 0xffffffff marks an empty slot).

template <class F, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_msm_accum(const Affine<F>* __restrict__ bases, const uint64_t* __restrict__ pairs,
            const uint32_t* __restrict__ d_count, uint32_t s_min, XYZZ<F>* buckets, XYZZ<F>* partial,
            uint32_t* partial_keys, uint32_t nthreads) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const uint32_t count = *d_count;
    uint32_t S = (count + nthreads - 1) / nthreads;
    if (S < s_min) S = s_min;
    const size_t lo = (size_t)t * S;
    uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
    if (lo < count) {
        const size_t hi = lo + S < count ? lo + S : count;
        uint32_t cur = (uint32_t)(pairs[lo] >> 32);
        const uint32_t first_key = cur;
        XYZZ<F> acc = XYZZ<F>::inf();
        bool first_done = false;
        for (size_t i = lo; i < hi; i++) {
            const uint64_t pr = pairs[i];
            uint32_t k = (uint32_t)(pr >> 32);
            if (k != cur) {
                if (!first_done) {
                    partial[2 * (size_t)t] = acc;
                    first_done = true;
                } else {
                    buckets[cur] = acc;
                }
                acc = XYZZ<F>::inf();
                cur = k;
            }
            uint32_t it = (uint32_t)pr;
            Affine<F> p = bases[it >> 1];
            acc.madd(p, it & 1);
        }
        k0 = first_key;
        if (!first_done) {
            partial[2 * (size_t)t] = acc;     // whole segment is one run
        } else {
            partial[2 * (size_t)t + 1] = acc;
            k1 = cur;
        }
    }
    partial_keys[2 * (size_t)t] = k0;
    partial_keys[2 * (size_t)t + 1] = k1;
}

// G2 over a field with the fused dual product: the same segment scheme with every Fq2 value spread over a lane pair
// (field.cuh Fp2Half).  Pair t owns segment t; both lanes walk the same sorted pairs (one broadcast load), each loads
// and stores its half of every coordinate (Fq2 = c0 || c1 in memory, so the halves of a coordinate are adjacent).
// Buckets and partial sums keep the XYZZ<Fq2> memory layout: the fold / reduce kernels are unchanged.
template <class F2, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_msm_accum_g2pair(const Affine<F2>* __restrict__ bases, const uint64_t* __restrict__ pairs,
                   const uint32_t* __restrict__ d_count, uint32_t s_min, XYZZ<F2>* buckets, XYZZ<F2>* partial,
                   uint32_t* partial_keys, uint32_t npairs) {
    using Fb = typename F2::base;
    using H = Fp2Half<Fb>;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t t = g >> 1, par = g & 1u;
    if (t >= npairs) return;
    const uint32_t count = *d_count;
    uint32_t S = (count + npairs - 1) / npairs;
    if (S < s_min) S = s_min;
    const size_t lo = (size_t)t * S;
    uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
    auto store = [par](XYZZ<F2>* dst, const XYZZ<H>& a) {
        Fb* d = reinterpret_cast<Fb*>(dst);
        d[0 + par] = a.x.v;
        d[2 + par] = a.y.v;
        d[4 + par] = a.zz.v;
        d[6 + par] = a.zzz.v;
    };
    if (lo < count) {
        const size_t hi = lo + S < count ? lo + S : count;
        uint32_t cur = (uint32_t)(pairs[lo] >> 32);
        const uint32_t first_key = cur;
        XYZZ<H> acc = XYZZ<H>::inf();
        bool first_done = false;
        // software pipeline: the base of item i + 1 is in flight while item i is added (254 registers, no spills;
        // -1 % at every size, which is what keeps this kernel level with the one-thread kernel on tables > 1 GB)
        uint64_t pr = pairs[lo];
        Affine<H> pn;
        {
            const Fb* src = reinterpret_cast<const Fb*>(bases + ((uint32_t)pr >> 1));
            pn.x.v = src[0 + par];
            pn.y.v = src[2 + par];
        }
        for (size_t i = lo; i < hi; i++) {
            const uint32_t k = (uint32_t)(pr >> 32);
            const uint32_t it = (uint32_t)pr;
            const Affine<H> p = pn;
            if (i + 1 < hi) {
                pr = pairs[i + 1];
                const Fb* src = reinterpret_cast<const Fb*>(bases + ((uint32_t)pr >> 1));
                pn.x.v = src[0 + par];
                pn.y.v = src[2 + par];
            }
            if (k != cur) {
                if (!first_done) {
                    store(&partial[2 * (size_t)t], acc);
                    first_done = true;
                } else {
                    store(&buckets[cur], acc);
                }
                acc = XYZZ<H>::inf();
                cur = k;
            }
            acc.madd(p, it & 1);
        }
        k0 = first_key;
        if (!first_done) {
            store(&partial[2 * (size_t)t], acc);     // whole segment is one run
        } else {
            store(&partial[2 * (size_t)t + 1], acc);
            k1 = cur;
        }
    }
    if (par == 0) {
        partial_keys[2 * (size_t)t] = k0;
        partial_keys[2 * (size_t)t + 1] = k1;
    }
}

// Level >= 1: same scheme over a list of XYZZ partials (full adds); empty slots skipped.
// `final_level`: single thread range covers everything -> all runs are complete.
// (A run that is cut by a high-order thread border stays split until the level whose segments swallow that border,
// so the chain runs all its levels: log2 of the list length, each a handful of dependent additions.)
template <class F, int S, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_msm_fold(const XYZZ<F>* __restrict__ in, const uint32_t* __restrict__ in_keys, uint32_t n_in,
           XYZZ<F>* buckets, XYZZ<F>* out, uint32_t* out_keys, uint32_t nthreads, int final_level) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    size_t lo = (size_t)t * S;
    size_t hi = lo + S < n_in ? lo + S : n_in;
    uint32_t cur = KEY_NONE, first_key = KEY_NONE;
    XYZZ<F> acc = XYZZ<F>::inf();
    bool first_done = false;
    for (size_t i = lo; i < hi; i++) {
        uint32_t k = in_keys[i];
        if (k == KEY_NONE) continue;
        if (cur == KEY_NONE) {
            cur = k;
            first_key = k;
        } else if (k != cur) {
            if (!first_done && !final_level) {
                out[2 * (size_t)t] = acc;
                first_done = true;
            } else {
                buckets[cur] = acc;
                first_done = true;
            }
            acc = XYZZ<F>::inf();
            cur = k;
        }
        acc.add(in[i]);
    }
    uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
    if (cur != KEY_NONE) {
        if (final_level) {
            buckets[cur] = acc;
        } else if (!first_done) {
            out[2 * (size_t)t] = acc;
            k0 = first_key;
        } else {
            out[2 * (size_t)t + 1] = acc;
            k0 = first_key;
            k1 = cur;
        }
    }
    if (!final_level) {
        out_keys[2 * (size_t)t] = k0;
        out_keys[2 * (size_t)t + 1] = k1;
    }
}

// ---- bucket reduction ------------------------------------------------------------------------
// chunk t of window w covers bucket values v = tG+1 .. tG+G (array index v-1):
//   sum v*B_v = sum (v - tG)*B_v + tG * sum B_v = acc + (tG)*run
template <class F>
__global__ void __launch_bounds__(128)
k_msm_reduce_chunks(const XYZZ<F>* __restrict__ buckets, uint32_t B, uint32_t G, uint32_t T, uint32_t W,
                    XYZZ<F>* chunk_out) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T * W) return;
    uint32_t w = g / T, t = g % T;
    const XYZZ<F>* bk = buckets + (size_t)w * B + (size_t)t * G;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    for (int j = (int)G - 1; j >= 0; j--) {
        XYZZ<F> b = bk[j];
        run.add(b);
        acc.add(run);
    }
    if (t != 0) {
        XYZZ<F> sc = run.mul_small(t * G);
        acc.add(sc);
    }
    chunk_out[g] = acc;
}

// Multi-block tree sum: every block adds 2*TPB consecutive inputs of window blockIdx.y.
template <class F, int TPB>
__global__ void __launch_bounds__(TPB)
k_msm_tree_sum(const XYZZ<F>* __restrict__ in, uint32_t m, uint32_t in_stride, XYZZ<F>* out, uint32_t out_stride) {
    extern __shared__ unsigned char smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    const XYZZ<F>* src = in + (size_t)blockIdx.y * in_stride;
    uint32_t i0 = blockIdx.x * 2 * TPB + threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::inf();
    if (i0 < m) acc = src[i0];
    if (i0 + TPB < m) acc.add(src[i0 + TPB]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = TPB / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            XYZZ<F> a = sh[threadIdx.x];
            a.add(sh[threadIdx.x + off]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)blockIdx.y * out_stride + blockIdx.x] = sh[0];
}

// ---- quad-cooperative variants of the latency-bound kernels (ec.cuh: quad_add / quad_dbl) -----------------
// Same lists, same outputs, four lanes per logical thread.  Used whenever the single-lane grid could not fill the GPU
// (late fold levels, the bucket reduction, the tree sums): there the kernel's duration is the longest dependent chain
// of point additions, and a quad walks it 3.3x faster.
template <class F, int S>
__global__ void __launch_bounds__(128)
k_msm_fold_quad(const XYZZ<F>* __restrict__ in, const uint32_t* __restrict__ in_keys, uint32_t n_in,
                XYZZ<F>* buckets, XYZZ<F>* out, uint32_t* out_keys, uint32_t nthreads, int final_level,
                const uint32_t* __restrict__ gate) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t t = g >> 2;
    const bool lead = (g & 3u) == 0;
    if (t >= nthreads) return;
    if (gate != nullptr && *gate == 0) return;     // k_msm_fold_resolve_quad has already finished the list
    size_t lo = (size_t)t * S;
    size_t hi = lo + S < n_in ? lo + S : n_in;
    uint32_t cur = KEY_NONE, first_key = KEY_NONE;
    XYZZ<F> acc = XYZZ<F>::inf();
    bool first_done = false;
    for (size_t i = lo; i < hi; i++) {
        uint32_t k = in_keys[i];
        if (k == KEY_NONE) continue;
        if (cur == KEY_NONE) {
            cur = k;
            first_key = k;
        } else if (k != cur) {
            if (!first_done && !final_level) {
                if (lead) out[2 * (size_t)t] = acc;
            } else {
                if (lead) buckets[cur] = acc;
            }
            first_done = true;
            acc = XYZZ<F>::inf();
            cur = k;
        }
        acc = quad_add(acc, in[i]);
    }
    uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
    if (cur != KEY_NONE) {
        if (final_level) {
            if (lead) buckets[cur] = acc;
        } else if (!first_done) {
            if (lead) out[2 * (size_t)t] = acc;
            k0 = first_key;
        } else {
            if (lead) out[2 * (size_t)t + 1] = acc;
            k0 = first_key;
            k1 = cur;
        }
    }
    if (!final_level && lead) {
        out_keys[2 * (size_t)t] = k0;
        out_keys[2 * (size_t)t + 1] = k1;
    }
}

// ---- all remaining borders at once ------------------------------------------------------------------------
// A partial list holds, per producer thread i, the sum of the first run of its range (slot 2i, "F_i") and of its last
// run (slot 2i+1, "L_i", absent when the whole range was one run); keys never decrease along the list.  The level-by-
// level fold pairs L_i with F_{i+1} only at the level whose segments swallow border i -- log2(threads) launches, each
// one dependent addition long.  Once a thread's range is longer than any bucket's run, every run of equal keys in the
// list has at most two pieces (L_i, F_{i+1}) and ALL borders can be closed in one launch: fold_runs classifies the (at
// most two) runs that start in thread i's slots; k_msm_fold_check raises a flag if any run has three or more pieces
// (a thread whose whole range lies inside one bucket's run: skewed scalars) or if the list has a hole before its end;
// k_msm_fold_resolve_quad closes every run when the flag is clear, and the remaining fold levels return at once
// (`gate`).  With the flag set nothing is written and the levels run as before.  fold_runs lives in fold_runs.cuh so
// that tests/test_fold_runs_host.py can fuzz it on the CPU against a brute-force grouping of the pieces.
static __global__ void k_msm_fold_check(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fold_runs(keys, i, n).bad) *flag = 1u;
}

template <class F>
__global__ void __launch_bounds__(128)
k_msm_fold_resolve_quad(const XYZZ<F>* __restrict__ in, const uint32_t* __restrict__ keys, uint32_t n,
                        XYZZ<F>* buckets, const uint32_t* __restrict__ flag) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = g >> 2;
    const bool lead = (g & 3u) == 0;
    if (i >= n || *flag != 0) return;
    const FoldRuns r = fold_runs(keys, i, n);
#pragma unroll 1
    for (int run = 0; run < 2; run++) {    // one quad_add call site: quads of a warp that diverge here still share it
        const bool have = run == 0 ? r.b : r.a;
        if (!have) continue;
        const bool merge = run == 0 ? r.b_merge : r.a_merge;
        XYZZ<F> acc = in[2 * (size_t)i + (run == 0 ? 1 : 0)];
        if (merge) acc = quad_add(acc, in[2 * (size_t)i + 2]);
        if (lead) buckets[run == 0 ? r.kb : r.ka] = acc;
    }
}

template <class F>
__global__ void __launch_bounds__(128)
k_msm_reduce_chunks_quad(const XYZZ<F>* __restrict__ buckets, uint32_t B, uint32_t G, uint32_t T, uint32_t W,
                         XYZZ<F>* chunk_out) {
    const uint32_t gq = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = gq >> 2;
    if (g >= T * W) return;
    uint32_t w = g / T, t = g % T;
    const XYZZ<F>* bk = buckets + (size_t)w * B + (size_t)t * G;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    for (int j = (int)G - 1; j >= 0; j--) {
        run = quad_add(run, bk[j]);
        acc = quad_add(acc, run);
    }
    if (t != 0) acc = quad_add(acc, quad_mul_small(run, t * G));
    if ((gq & 3u) == 0) chunk_out[g] = acc;
}

// every block adds 2 * TPB consecutive inputs of set blockIdx.y with TPB quads (blockDim.x = 4 * TPB)
template <class F, int TPB>
__global__ void __launch_bounds__(4 * TPB)
k_msm_tree_sum_quad(const XYZZ<F>* __restrict__ in, uint32_t m, uint32_t in_stride, XYZZ<F>* out, uint32_t out_stride) {
    extern __shared__ unsigned char smem_raw[];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(smem_raw);
    const XYZZ<F>* src = in + (size_t)blockIdx.y * in_stride;
    const uint32_t qi = threadIdx.x >> 2;
    const bool lead = (threadIdx.x & 3u) == 0;
    const uint32_t i0 = blockIdx.x * 2 * TPB + qi;
    XYZZ<F> acc = XYZZ<F>::inf();
    if (i0 < m) acc = src[i0];
    if (i0 + TPB < m) acc = quad_add(acc, src[i0 + TPB]);
    if (lead) sh[qi] = acc;
    __syncthreads();
    for (int off = TPB / 2; off > 0; off >>= 1) {
        XYZZ<F> a = XYZZ<F>::inf();
        const bool work = (int)qi < off;
        if (work) a = quad_add(sh[qi], sh[qi + off]);
        __syncthreads();                       // every quad has read its two inputs
        if (work && lead) sh[qi] = a;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)blockIdx.y * out_stride + blockIdx.x] = sh[0];
}

// tbl[w*n + i] = 2^(c w) * P_i, affine (one inversion per entry; one-off per base set)
template <class F>
__global__ void __launch_bounds__(128)
k_msm_precompute(const Affine<F>* __restrict__ pts, size_t n, int c, int W, Affine<F>* tbl, uint8_t* skip) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = pts[i];
    tbl[i] = p;
    skip[i] = p.is_inf() ? 1 : 0;
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        tbl[(size_t)w * n + i] = acc.to_affine();
    }
}

// one block per scalar vector: Horner over its W window sums (W = 1 in table mode: no doublings)
template <class F>
__global__ void k_msm_combine(const XYZZ<F>* __restrict__ wsum_all, int W, int c, Jacobian<F>* out) {
    if (threadIdx.x >= 4) return;              // one quad per scalar vector (ec.cuh quad_dbl / quad_add)
    const XYZZ<F>* wsum = wsum_all + (size_t)blockIdx.x * W;
    XYZZ<F> acc = wsum[W - 1];
    for (int w = W - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = quad_dbl(acc);
        acc = quad_add(acc, wsum[w]);
    }
    if (threadIdx.x == 0) out[blockIdx.x] = Jacobian<F>::from_xyzz(acc);
}

template <class F>
__global__ void k_set_identity_jac(Jacobian<F>* out, size_t K) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < K) out[i] = {F::one(), F::one(), F::zero()};
}

template <class Curve, int GROUP>
void bases_precompute(hk_ctx* ctx, BaseSet& bs, double density) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    // one builder at a time per base set; the table pointer is published last (forked contexts share base sets)
    std::lock_guard<std::mutex> lk(bs.tbl_mu);
    if (bs.tbl || bs.n == 0) return;
    MsmPlan pl = msm_plan(bs.n, Curve::FR_BITS, true, density);
    HK_REQUIRE((size_t)pl.W * bs.n < (size_t(1) << 31), HK_ERR_UNSUPPORTED, "precompute: windows * n must be < 2^31");
    void* t = nullptr;
    uint8_t* sk = nullptr;
    cudaError_t e = cudaMalloc(&t, (size_t)pl.W * bs.n * sizeof(Affine<F>) + 256);
    if (e == cudaSuccess) e = cudaMalloc((void**)&sk, bs.n + 256);
    if (e != cudaSuccess) {
        if (t) cudaFree(t);
        throw HkError{HK_ERR_NOMEM, std::string("cudaMalloc(window table): ") + cudaGetErrorString(e)};
    }
    k_msm_precompute<F><<<ceil_div(bs.n, 128), 128, 0, ctx->stream>>>(reinterpret_cast<const Affine<F>*>(bs.d), bs.n, pl.c,
                                                                     pl.W, reinterpret_cast<Affine<F>*>(t), sk);
    ctx->launches++;
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) le = cudaStreamSynchronize(ctx->stream);
    if (le != cudaSuccess) {
        cudaFree(t);
        cudaFree(sk);
        throw HkError{HK_ERR_CUDA, std::string("k_msm_precompute: ") + cudaGetErrorString(le)};
    }
    bs.skip = sk;
    bs.tbl_c = pl.c;
    bs.tbl_W = pl.W;
    bs.tbl = t;
}

// Steps 1-3: key-sorted digit list of K scalar vectors for a base range of `n` points.  Runs on `lane`'s control
// stream and records ev_sorted[lane]; the list lives in that lane's workspace until the lane sorts again.
// Only the GEOMETRY of the base set matters (table or not, window size, infinity flags), so base sets that agree
// in these can share one list (msm_accumulate with a different MsmBases).
template <class Curve>
MsmSorted msm_sort(hk_ctx* ctx, const MsmBases& bv, size_t offset, size_t n, const Scalar256* d_scalars_in, size_t K,
                   bool scalars_mont, size_t sc_stride, int lane) {
    using Fr = typename Curve::Fr;
    if (lane < 0 || lane >= hk_ctx::LANES) lane = 0;
    cudaStream_t st = ctx->lane_stream(lane);
    MsmWs& ws = ctx->mws[lane];
    MsmSorted so;
    so.lane = lane;
    so.n = n;
    so.K = K;
    so.offset = offset;
    if (sc_stride == 0) sc_stride = n;
    HK_REQUIRE(n < (size_t(1) << 31), HK_ERR_UNSUPPORTED, "msm: n must be < 2^31");
    so.table = bv.tbl != nullptr;
    so.pl = so.table ? msm_plan_for(bv.c, Curve::FR_BITS) : msm_plan(n, Curve::FR_BITS, false);
    if (so.table) HK_REQUIRE(so.pl.W == bv.W, HK_ERR_INVALID, "msm: window table geometry mismatch");
    const MsmPlan& pl = so.pl;
    const uint32_t per_window = so.table ? 0u : 1u;
    so.item_stride = so.table ? (uint32_t)bv.n_total : 0u;
    so.n_total = bv.n_total;
    so.Wred = so.table ? 1u : (uint32_t)pl.W;        // bucket sets per scalar vector
    so.nsets = (uint32_t)K * so.Wred;
    so.nkeys = (size_t)so.nsets * pl.B;
    so.max_items = n * K * (size_t)pl.W;
    HK_REQUIRE(so.nkeys < (size_t(1) << 31), HK_ERR_UNSUPPORTED, "msm: too many buckets for one batch");
    HK_REQUIRE(so.max_items < (size_t(1) << 32), HK_ERR_UNSUPPORTED, "msm: batch * n * windows must be < 2^32");
    const Scalar256* d_sc = d_scalars_in;
    if (scalars_mont) {
        // convert the used part of every vector into a dense K x n array
        Scalar256* conv = ws.scalars.as<Scalar256>(n * K);
        if (sc_stride == n) {
            k_scalars_from_mont<Fr><<<ceil_div(n * K, 256), 256, 0, st>>>(d_scalars_in, conv, n * K);
            HK_LAUNCH_CHECK(ctx);
        } else {
            for (size_t k = 0; k < K; k++) {
                k_scalars_from_mont<Fr><<<ceil_div(n, 256), 256, 0, st>>>(d_scalars_in + k * sc_stride, conv + k * n, n);
                HK_LAUNCH_CHECK(ctx);
            }
        }
        d_sc = conv;
        sc_stride = n;
    }
    uint64_t* pairs = ws.items.as<uint64_t>(so.max_items + 2);
    const uint8_t* skip = bv.skip ? bv.skip + offset : nullptr;
    // blocks per vector for the shared-memory sort: about two resident waves (one block per SM when the counters
    // take more than half of the shared memory), never less than 1024 scalars per block
    const size_t smem_bytes = (size_t)pl.B * sizeof(uint32_t);
    uint32_t J = 0;
    if (so.table && smem_bytes <= 160 * 1024 && !ctx->opt_sort_global) {
        // about 2 x 148 blocks when a block takes a whole SM (measured at 2^20: hist 100 / scan 83 / scatter 309 us
        // with two waves, 67 / 48 / 443 us with one: the scatter wants the parallelism), 4 x 148 small ones otherwise
        const uint32_t per_sm = smem_bytes > 100 * 1024 ? 2u : 4u;
        uint64_t want = ((uint64_t)ctx->sm_count * per_sm + K - 1) / K;
        const uint64_t most = (n + 1023) / 1024;
        if (want > most) want = most;
        if (want < 1) want = 1;
        J = (uint32_t)want;
        if ((uint64_t)so.nkeys * J >= (uint64_t(1) << 31)) J = 0;
    }
    uint32_t* d_count;
    if (J) {
        const uint32_t chunk = (uint32_t)((n + J - 1) / J);
        const size_t ncnt = so.nkeys * J;
        uint32_t* cnt = ws.hist.as<uint32_t>(ncnt + 4);
        d_count = cnt + ncnt;
        static bool attr_set[64] = {};     // opt in to > 48 KB of dynamic shared memory once per device
        if (!attr_set[ctx->device & 63]) {
            HK_CUDA(cudaFuncSetAttribute(k_msm_sort_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            HK_CUDA(cudaFuncSetAttribute(k_msm_sort_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr_set[ctx->device & 63] = true;
        }
        const dim3 grid(J, (unsigned)K);
        k_msm_sort_smem<false><<<grid, 1024, smem_bytes, st>>>(d_sc, n, sc_stride, pl.c, pl.W, pl.B, J, chunk, so.item_stride,
                                                              skip, cnt, pairs);
        HK_LAUNCH_CHECK(ctx);
        exclusive_scan_u32(ctx, ws, st, cnt, ncnt, d_count);
        k_msm_sort_smem<true><<<grid, 1024, smem_bytes, st>>>(d_sc, n, sc_stride, pl.c, pl.W, pl.B, J, chunk, so.item_stride,
                                                             skip, cnt, pairs);
        HK_LAUNCH_CHECK(ctx);
    } else {
        uint32_t* hist = ws.hist.as<uint32_t>(so.nkeys + 4);
        d_count = hist + so.nkeys;     // total item count lives right after the histogram
        HK_CUDA(cudaMemsetAsync(hist, 0, (so.nkeys + 4) * sizeof(uint32_t), st));
        const unsigned sort_blocks = ceil_div(n * K, 256);
        k_msm_hist<<<sort_blocks, 256, 0, st>>>(d_sc, n, K, sc_stride, pl.c, pl.W, pl.B, so.Wred, per_window, skip, hist);
        HK_LAUNCH_CHECK(ctx);
        exclusive_scan_u32(ctx, ws, st, hist, so.nkeys, d_count);
        k_msm_scatter<<<sort_blocks, 256, 0, st>>>(d_sc, n, K, sc_stride, pl.c, pl.W, pl.B, so.Wred, per_window,
                                                   so.item_stride, skip, hist, pairs);
        HK_LAUNCH_CHECK(ctx);
    }
    HK_CUDA(cudaEventRecord(ctx->ev_sorted[lane], st));
    so.pairs = pairs;
    so.d_count = d_count;
    return so;
}

// Steps 4-7 over a sorted list: accumulate the bases of `bv` (same geometry as the set the list was sorted for),
// fold, reduce, combine -> K Jacobian points.  Runs on `lane`'s streams (bulk stream for the accumulation) with
// the lane's own bucket / partial scratch; waits for the sort through ev_sorted[so.lane].
template <class Curve, int GROUP>
void msm_accumulate(hk_ctx* ctx, const MsmSorted& so, const MsmBases& bv, void* d_out_jac_v, int lane) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    Jacobian<F>* d_out = reinterpret_cast<Jacobian<F>*>(d_out_jac_v);
    if (lane < 0 || lane >= hk_ctx::LANES) lane = 0;
    cudaStream_t st = ctx->lane_stream(lane);      // control stream: folds, reduction
    cudaStream_t sb = ctx->bulk_stream(lane);      // the accumulation kernel (== st on lane 0)
    MsmWs& ws = ctx->mws[lane];
    const MsmPlan& pl = so.pl;
    const size_t K = so.K;
    HK_REQUIRE((bv.tbl != nullptr) == so.table && (!so.table || (bv.c == pl.c && bv.W == pl.W && bv.n_total == so.n_total)),
               HK_ERR_INVALID, "msm: base set geometry differs from the one the list was sorted for");
    const Affine<F>* d_bases = reinterpret_cast<const Affine<F>*>(so.table ? bv.tbl : bv.pts) + so.offset;
    const uint32_t nsets = so.nsets;
    const size_t nkeys = so.nkeys;
    // accumulation geometry: a fixed number of waves of resident CTAs; the segment length follows from the item
    // count on the device (never below s_min).  G1: 4 CTAs/SM at 128 registers, G2: 2 CTAs/SM (k_msm_accum MINB).
    constexpr int MINB = GROUP == 2 ? 1 : 4;
    const uint32_t resident = (uint32_t)ctx->sm_count * (GROUP == 2 ? 2u : 4u) * 128u;
    // a standalone MSM whose list cannot fill the GPU with 32-item segments (latency-bound chains) uses 8
    uint32_t s_min = ctx->opt_accum_smin ? ctx->opt_accum_smin : ((lane == 0 && so.max_items <= (size_t(1) << 21)) ? 8u : 32u);
    // waves = 0 (default): one thread per s_min items (fine-grained CTAs balance the tail best: measured
    // profiles/r02b -- 4 waves cost 2.5 % of the accumulation, which is what the smaller fold saves)
    const uint32_t waves = ctx->opt_accum_waves;
    uint64_t want = (so.max_items + s_min - 1) / s_min;
    if (waves && want > (uint64_t)waves * resident) want = (uint64_t)waves * resident;
    if (want < 1) want = 1;
    const uint32_t nthr0 = (uint32_t)want;
    // fold segment length (levels that remain after the gather step): 16 inside a proof batch (few launches; the
    // tails hide under other lanes' accumulations), 4 for a standalone MSM (short serial chains)
    const int S1 = ctx->opt_fold_seg ? (int)ctx->opt_fold_seg : (lane == 0 ? 4 : 16);
    // Launches of at most this many logical threads cannot fill the GPU (148 SMs x 256 resident fold threads): they
    // are latency-bound and, in a STANDALONE MSM (lane 0), run quad-cooperatively -- 2^16: 1.43 -> 1.07 ms (G1),
    // 4.06 -> 2.85 ms (G2).  Inside a proof batch the tails hide under other lanes' accumulations and the quad
    // kernels' ~20 % extra multiplier work costs throughput (286 vs 287 proofs/s), so the lanes keep one thread per
    // addition.  (HK_OPT_QUAD_MAX_THREADS overrides the limit for every lane; 1 = never.)
    const size_t quad_max = ctx->opt_quad_max ? ctx->opt_quad_max : (lane == 0 ? 32768 : 1);
    // (Measured and rejected, profiles/r02_summary.md: 4-entry segments for the first three levels and 16-entry ones
    // for the mostly empty late levels -- 9 instead of 19 launches at 2^20 -- is SLOWER, 7.55 vs 7.31 ms: a late level
    // costs the latency of its longest thread, and 16-entry threads are longer.)
    std::vector<uint32_t> lvl_in;      // entries consumed at fold level l
    {
        uint32_t cur = 2 * nthr0;
        while (true) {
            lvl_in.push_back(cur);
            uint32_t thr = ceil_div(cur, S1);
            if (thr <= 1) break;
            cur = 2 * thr;
        }
    }
    size_t part_total = 0;
    for (uint32_t v : lvl_in) part_total += v;
    XYZZ<F>* buckets = ws.buckets.as<XYZZ<F>>(nkeys + 1);      // + the fold's "runs left" flag, cleared with the buckets
    uint32_t* d_fold_flag = reinterpret_cast<uint32_t*>(buckets + nkeys);
    XYZZ<F>* partial = ws.chunks.as<XYZZ<F>>(part_total);
    uint32_t* partial_keys = ws.misc.as<uint32_t>(part_total);
    // reduction geometry: short serial chunks, then a multi-block tree
    // standalone (lane 0): short chunks shorten the serial tail; inside a proof batch the larger grids of a
    // latency-bound kernel take SM slots from the accumulations, so the lanes keep long chunks
    // (lane 0 with the quad-cooperative kernels: 4 buckets per quad from B = 8192 on -- 2^20: 7.10 -> 6.95 ms, 2^18:
    // 2.54 -> 2.37 ms; below that 2 -- the scalar multiplication of the chunk offset dominates a chunk's chain)
    uint32_t G = lane == 0 ? (pl.B >= 8192 ? 4 : (pl.B >= 64 ? 2 : 1))
                           : (pl.B >= 32768 ? 8 : (pl.B >= 1024 ? 4 : (pl.B >= 64 ? 2 : 1)));
    if (ctx->opt_reduce_g > 0 && ctx->opt_reduce_g <= pl.B) G = ctx->opt_reduce_g;
    uint32_t T = pl.B / G;
    constexpr int TS_TPB = 64;
    static_assert(TS_TPB * sizeof(XYZZ<F>) <= 48 * 1024, "tree-sum tile must fit the default dynamic smem limit");
    std::vector<uint32_t> tree;        // tree levels: T -> ceil(T / 128) -> ... -> 1 (per bucket set)
    for (uint32_t m = T; m > 1; m = ceil_div(m, 2 * TS_TPB)) tree.push_back(m);
    size_t tree_total = (size_t)T * nsets;
    for (uint32_t m : tree) tree_total += (size_t)ceil_div(m, 2 * TS_TPB) * nsets;
    XYZZ<F>* chunk_out = ws.wsum.as<XYZZ<F>>(tree_total + nsets);

    HK_CUDA(cudaMemsetAsync(buckets, 0, (nkeys + 1) * sizeof(XYZZ<F>), st));
    // the accumulation needs the sorted list (possibly another lane's) and this lane's cleared buckets
    if (sb != st) {
        HK_CUDA(cudaEventRecord(ctx->ev_ready[lane], st));
        HK_CUDA(cudaStreamWaitEvent(sb, ctx->ev_ready[lane], 0));
    }
    if (sb != ctx->lane_stream(so.lane)) HK_CUDA(cudaStreamWaitEvent(sb, ctx->ev_sorted[so.lane], 0));
    int pi = prof_begin(ctx, GROUP == 1 ? HK_PROF_MSM_ACCUM_G1 : HK_PROF_MSM_ACCUM_G2, sb);
    bool launched = false;
#if defined(HK_MUL2) && !defined(HK_NO_G2_PAIR)
    // An Fq2 point addition split over a lane pair (field.cuh Fp2Half): bit-identical, no spills (the one-thread kernel:
    // 255 registers with 1 KB of spills), multiplier pipe 81.9 % vs 72.1 % busy in the `ncu --set full` captures at
    // 2^18 (profiles/r02p_*): 4.64 vs 5.16 ms.  Standalone G2 MSM, ms, pair vs one thread: 2^16 2.57 / 2.67,
    // 2^18 6.11 / 6.71, 2^20 20.79 / 20.79, 2^22 73.3 / 74.1; proof pipeline +1 %.  Default for every field with the
    // fused dual product (BLS12-381); HK_OPT_G2_LANE_PAIR = 2 selects the one-thread kernel.
    if constexpr (GROUP == 2) {
        if constexpr (F::base::HAS_MUL2) {
            if (ctx->opt_g2_pair != 2) {
                k_msm_accum_g2pair<F, 2><<<ceil_div(2 * (size_t)nthr0, 128), 128, 0, sb>>>(d_bases, so.pairs, so.d_count, s_min,
                                                                                        buckets, partial, partial_keys, nthr0);
                launched = true;
            }
        }
    }
#endif
    if (!launched)
        k_msm_accum<F, MINB><<<ceil_div(nthr0, 128), 128, 0, sb>>>(d_bases, so.pairs, so.d_count, s_min, buckets, partial,
                                                                   partial_keys, nthr0);
    HK_LAUNCH_CHECK(ctx);
    prof_end(ctx, pi, 0, so.d_count, sb);
    if (sb != st) {
        HK_CUDA(cudaEventRecord(ctx->ev_accum[lane], sb));
        HK_CUDA(cudaStreamWaitEvent(st, ctx->ev_accum[lane], 0));
    }
    // Where a thread's range (s_min items, times S1 / 2 per level) has grown past 1.5 x the longest run expected of a
    // bucket and the list is short enough for the quad kernels, all borders are closed at once (k_msm_fold_resolve_quad)
    // and the remaining levels are gated off; skewed inputs fail the check and fold level by level as before.
    // Standalone G1 2^16: 11 launches of ~12.5 us -> 2 + 11 empty ones.
    // expected longest run of a bucket for uniform scalars: n / 2^(c-1) items per full window, and the top window
    // (t = bits - c (W - 1) bits wide) piles its n digits onto 2^(t-1) buckets; all windows share the buckets of a table
    const double n_vec = double(so.max_items) / double(K * (size_t)pl.W);
    const int top_bits = (int)Curve::FR_BITS - pl.c * (pl.W - 1);
    const double run_full = n_vec / double(pl.B), run_top = n_vec / double(1u << (top_bits > 1 ? (top_bits < pl.c ? top_bits - 1 : pl.c - 1) : 0));
    const double long_run = so.table ? (pl.W - 1) * run_full + run_top : (run_full > run_top ? run_full : run_top);
    double range = double(s_min);
    bool resolved = false;
    size_t off = 0;
    for (size_t l = 0; l < lvl_in.size(); l++) {
        uint32_t n_in = lvl_in[l];
        uint32_t thr = ceil_div(n_in, S1);
        int final_level = (l + 1 == lvl_in.size());
        XYZZ<F>* in = partial + off;
        uint32_t* in_keys = partial_keys + off;
        XYZZ<F>* out = partial + off + n_in;
        uint32_t* out_keys = partial_keys + off + n_in;
        if (!resolved && !final_level && !ctx->opt_no_fold_resolve && n_in / 2 <= quad_max && thr <= quad_max &&
            range >= 1.5 * long_run) {
            const uint32_t n = n_in / 2;
            k_msm_fold_check<<<ceil_div(n, 256), 256, 0, st>>>(in_keys, n, d_fold_flag);
            HK_LAUNCH_CHECK(ctx);
            k_msm_fold_resolve_quad<F><<<ceil_div(4 * (size_t)n, 128), 128, 0, st>>>(in, in_keys, n, buckets, d_fold_flag);
            HK_LAUNCH_CHECK(ctx);
            resolved = true;
        }
        range *= double(S1 / 2);
        const uint32_t* gate = resolved ? d_fold_flag : nullptr;
#define HK_FOLD(SS)                                                                                                      \
    do {                                                                                                                \
        if (thr <= quad_max)                                                                                            \
            k_msm_fold_quad<F, SS><<<ceil_div(4 * (size_t)thr, 128), 128, 0, st>>>(in, in_keys, n_in, buckets, out, out_keys, thr, final_level, gate); \
        else                                                                                                            \
            k_msm_fold<F, SS, 1><<<ceil_div(thr, 128), 128, 0, st>>>(in, in_keys, n_in, buckets, out, out_keys, thr, final_level); \
    } while (0)
        if (S1 == 8) HK_FOLD(8);
        else if (S1 == 4) HK_FOLD(4);
        else if (S1 == 32) HK_FOLD(32);
        else HK_FOLD(16);
#undef HK_FOLD
        HK_LAUNCH_CHECK(ctx);
        off += n_in;
    }
    if ((size_t)T * nsets <= quad_max)
        k_msm_reduce_chunks_quad<F><<<ceil_div(4 * (size_t)T * nsets, 128), 128, 0, st>>>(buckets, pl.B, G, T, nsets, chunk_out);
    else
        k_msm_reduce_chunks<F><<<ceil_div((size_t)T * nsets, 128), 128, 0, st>>>(buckets, pl.B, G, T, nsets, chunk_out);
    HK_LAUNCH_CHECK(ctx);
    XYZZ<F>* cur = chunk_out;
    uint32_t cur_stride = T;
    XYZZ<F>* nxt = chunk_out + (size_t)T * nsets;
    for (uint32_t m : tree) {
        uint32_t nb = ceil_div(m, 2 * TS_TPB);
        if ((size_t)nb * nsets * TS_TPB <= quad_max)
            k_msm_tree_sum_quad<F, TS_TPB><<<dim3(nb, nsets), 4 * TS_TPB, TS_TPB * sizeof(XYZZ<F>), st>>>(cur, m, cur_stride, nxt, nb);
        else
            k_msm_tree_sum<F, TS_TPB><<<dim3(nb, nsets), TS_TPB, TS_TPB * sizeof(XYZZ<F>), st>>>(cur, m, cur_stride, nxt, nb);
        HK_LAUNCH_CHECK(ctx);
        cur = nxt;
        cur_stride = nb;
        nxt += (size_t)nb * nsets;
    }
    // cur: one point per bucket set, K * Wred of them, vector-major
    k_msm_combine<F><<<(unsigned)K, 32, 0, st>>>(cur, (int)so.Wred, pl.c, d_out);
    HK_LAUNCH_CHECK(ctx);
}

template <class Curve, int GROUP>
void msm_run(hk_ctx* ctx, const MsmBases& bv, size_t offset, size_t n, const Scalar256* d_scalars_in, size_t K,
             bool scalars_mont, void* d_out_jac_v, size_t sc_stride, int lane) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    if (lane < 0 || lane >= hk_ctx::LANES) lane = 0;
    if (K == 0) return;
    if (n == 0) {
        k_set_identity_jac<F><<<ceil_div(K, 128), 128, 0, ctx->lane_stream(lane)>>>(reinterpret_cast<Jacobian<F>*>(d_out_jac_v), K);
        HK_LAUNCH_CHECK(ctx);
        return;
    }
    MsmSorted so = msm_sort<Curve>(ctx, bv, offset, n, d_scalars_in, K, scalars_mont, sc_stride, lane);
    msm_accumulate<Curve, GROUP>(ctx, so, bv, d_out_jac_v, lane);
}

}  // namespace hk
