// Pippenger MSM for sm_100a.  Replaces ark-ec 0.4 `VariableBaseMSM::msm_bigint` as called from
// /root/reference/cp-groth16/src/prover.rs:117,129,168 and committer.rs:89,113.
//
// Pipeline (all on the context stream, no host synchronisation):
//   1. k_msm_hist      signed-digit recoding of every scalar (window c, digits in
//                      [-2^(c-1), 2^(c-1)], top window unrecoded) + histogram of
//                      (window, |digit|) keys
//   2. exclusive scan  bucket start offsets
//   3. k_msm_scatter   counting sort: item = (point index << 1 | sign) placed in its bucket run
//   4. k_msm_accum     balanced segmented accumulation: every thread owns a fixed-length
//                      segment of the sorted item list (so skewed 0/1-heavy witness scalars
//                      cost the same as uniform ones), interior runs go straight to their
//                      bucket, the two boundary runs of each segment go to a partial list
//                      that is folded by k_msm_fold (recursively, same scheme)
//   5. k_msm_reduce*   per window sum_b b*B_b by chunked running sums, then a multi-block tree sum
//   6. k_msm_combine   Horner over windows -> one Jacobian point
//
// Table mode (base sets with a window-shifted table, tbl[w*n+i] = 2^(c w) P_i): every window's
// digit selects from its own table row, so ALL windows share ONE bucket set: keys are bucket
// values only, the reduce stage is W times smaller and step 6 has no doublings at all -- the
// serial tail that dominated capture r01a disappears.  Memory cost: W x the base set (HBM is 180 GB).
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace hk {

struct MsmPlan {
    int c;       // window bits
    int W;       // windows
    uint32_t B;  // buckets per window = 2^(c-1)
};

inline MsmPlan msm_plan_for(int c, int scalar_bits) {
    return {c, (scalar_bits + 1 + c - 1) / c, 1u << (c - 1)};
}

// Window size from a cost model in Fq multiplications, calibrated on B200 (profiles/r01b):
//   accumulate: 10 per item; items = density * n per window (+ ~n/2 once for 0/1 scalars when
//               density < 1: a scalar equal to 1 has a single non-zero digit)
//   reduce:     ~300 per bucket (chunked running sums + chunk offset multiply + tree; these
//               kernels are latency-bound, far from the multiplier peak), per window without a
//               table, once with a table
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
        carry = 0;
        return (int)coef;
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
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    bool active = g < n * K;
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
    if (!__any_sync(0xffffffffu, active)) return;
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

// pairs[pos] = key << 32 | item: the accumulation kernel reads one 8-byte stream instead of two 4-byte ones
static __global__ void k_msm_scatter(const Scalar256* sc, size_t n, size_t K, size_t sc_stride, int c, int W, uint32_t B,
                                     uint32_t sets_per_vec, uint32_t per_window, uint32_t item_stride,
                                     const uint8_t* __restrict__ skip, uint32_t* cursor, uint64_t* pairs) {
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    bool active = g < n * K;
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
    if (!__any_sync(0xffffffffu, active)) return;
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
// Level 0: thread t owns items [t*S, (t+1)*S) of the key-sorted list (mixed adds of affine
// bases).  Runs strictly inside the segment are complete -> bucket array.  The first and the
// last run of a segment may continue in a neighbour -> partial list slots 2t, 2t+1
// (key = 0xffffffff marks an empty slot).
static constexpr uint32_t KEY_NONE = 0xffffffffu;

// SENT: the list comes from the batch-affine rounds (padding slots hold KEY_NONE, length = *d_count >> shift).
template <class F, int S, int MINB, bool SENT>
__global__ void __launch_bounds__(128, MINB)
k_msm_accum(const Affine<F>* __restrict__ bases, const uint64_t* __restrict__ pairs,
            const uint32_t* __restrict__ d_count, uint32_t count_shift, XYZZ<F>* buckets, XYZZ<F>* partial,
            uint32_t* partial_keys, uint32_t nthreads) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    size_t lo = (size_t)t * S;
    uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
    if constexpr (!SENT) {
        uint32_t count = *d_count;
        if (lo < count) {
            size_t hi = lo + S < count ? lo + S : count;
            uint32_t cur = (uint32_t)(pairs[lo] >> 32);
            uint32_t first_key = cur;
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
    } else {
        const uint32_t count = *d_count >> count_shift;
        if (lo < count) {
            size_t hi = lo + S < count ? lo + S : count;
            uint32_t cur = KEY_NONE, first_key = KEY_NONE;
            XYZZ<F> acc = XYZZ<F>::inf();
            bool first_done = false;
            for (size_t i = lo; i < hi; i++) {
                const uint64_t pr = pairs[i];
                uint32_t k = (uint32_t)(pr >> 32);
                if (k == KEY_NONE) continue;                  // padding slot
                if (cur == KEY_NONE) {
                    cur = k;
                    first_key = k;
                } else if (k != cur) {
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
            if (cur != KEY_NONE) {
                k0 = first_key;
                if (!first_done) {
                    partial[2 * (size_t)t] = acc;
                } else {
                    partial[2 * (size_t)t + 1] = acc;
                    k1 = cur;
                }
            }
        }
    }
    partial_keys[2 * (size_t)t] = k0;
    partial_keys[2 * (size_t)t + 1] = k1;
}

// Level >= 1: same scheme over a list of XYZZ partials (full adds); empty slots skipped.
// `final_level`: single thread range covers everything -> all runs are complete.
template <class F, int S>
__global__ void __launch_bounds__(128)
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

// ---- batch-affine pre-reduction ---------------------------------------------------------------
// Before the XYZZ accumulation the key-sorted list is halved R times: in every round adjacent entries of the
// same bucket are added pairwise IN AFFINE coordinates (lambda = (y2 - y1) / (x2 - x1), 2M + 1S) with ONE
// shared inversion per round for the whole launch (Montgomery's trick: 3 more multiplications per addition),
// i.e. 5M + 1S (+1M to recompute the prefix) instead of the 8M + 2S of a mixed XYZZ addition.  An inversion
// costs 500+ multiplications, so it only pays when shared grid-wide:
//   k_ba_prefix   every thread owns S consecutive slots, multiplies the denominators of its pairs -> tot[t]
//   ba_invert     tot[] -> 1 / tot[] by a 64-ary product tree (forward), one GCD inversion per top-level
//                 entry (a warp each, lane 0), and the tree backwards
//   k_ba_apply    every thread recomputes its prefix products, walks its pairs backwards peeling off the
//                 individual inverses, adds, and writes the round's output (odd leftovers pass through)
// Layout: with the rounds enabled every bucket's run of the sorted list is padded to a multiple of 2^R slots
// (the histogram counts are rounded up before the scan; unused slots hold the KEY_NONE sentinel).  Every run
// then starts at a multiple of 2^R, so in EVERY round slot 2j pairs with slot 2j+1 and their sum lands in slot
// j of the next list: no per-round offsets, no lookups, and all lanes of a warp take the same branch.
// After R rounds (7/8 of the additions of long runs) the remaining list goes through k_msm_accum unchanged,
// with the last round's points as its bases.
enum : int { BA_ADD = 0, BA_DBL = 1, BA_INF = 2, BA_TAKE1 = 3, BA_TAKE2 = 4 };
static constexpr uint64_t BA_EMPTY = ~0ull;      // pairs[] sentinel: key = KEY_NONE

// classify p1 + p2 and return the denominator that the shared inversion has to cover (1 when none is needed)
template <class F>
__device__ __forceinline__ int ba_classify(const Affine<F>& p1, const Affine<F>& p2, F& d) {
    d = F::one();
    if (p1.is_inf()) return BA_TAKE2;
    if (p2.is_inf()) return BA_TAKE1;
    F dx = p2.x - p1.x;
    if (!dx.is_zero()) {
        d = dx;
        return BA_ADD;
    }
    if (p1.y == p2.y) {          // same point: tangent, denominator 2y (y != 0: no 2-torsion in these groups)
        d = p1.y.dbl();
        return BA_DBL;
    }
    return BA_INF;               // p2 = -p1
}

template <class F>
__device__ __forceinline__ Affine<F> ba_add(const Affine<F>& p1, const Affine<F>& p2, const F& inv_d, int kind) {
    if (kind == BA_TAKE1) return p1;
    if (kind == BA_TAKE2) return p2;
    if (kind == BA_INF) return Affine<F>::inf();
    F lam;
    if (kind == BA_DBL) {
        F x2 = p1.x.sqr();
        lam = (x2.dbl() + x2) * inv_d;
    } else {
        lam = (p2.y - p1.y) * inv_d;
    }
    F x3 = lam.sqr() - p1.x - p2.x;
    F y3 = lam * (p1.x - x3) - p1.y;
    return {x3, y3};
}

template <class F>
__device__ __forceinline__ Affine<F> ba_load(const Affine<F>* __restrict__ pts, uint32_t item) {
    Affine<F> p = pts[item >> 1];
    if ((item & 1u) && !p.is_inf()) p.y = p.y.neg();
    return p;
}

// hist[k] = cnt rounded up to a multiple of 2^R (before the scan)
static __global__ void k_ba_pad_counts(uint32_t* hist, size_t nkeys, uint32_t mask) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < nkeys) hist[k] = (hist[k] + mask) & ~mask;
}
static __global__ void k_ba_fill_empty(uint64_t* pairs, const uint32_t* __restrict__ d_n) {
    const size_t n = *d_n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        pairs[i] = BA_EMPTY;
}

// Denominator of pair (p1, p2) from as little data as possible: the x coordinates decide almost always.
template <class F>
__device__ __forceinline__ F ba_denominator(const Affine<F>* __restrict__ pts, uint32_t it1, uint32_t it2) {
    const Affine<F>* a = pts + (it1 >> 1);
    const Affine<F>* b = pts + (it2 >> 1);
    F x1 = a->x, x2 = b->x;
    F dx = x2 - x1;
    if (!dx.is_zero() && !x1.is_zero() && !x2.is_zero()) return dx;      // generic addition
    // rare: equal x (doubling or P - P) or a possible point at infinity (x = 0): decide on the full points
    Affine<F> p1 = ba_load(pts, it1), p2 = ba_load(pts, it2);
    F d;
    ba_classify(p1, p2, d);
    return d;
}

// n slots this round = *d_n >> shift (the padded total of the sort, halved once per earlier round).
// Thread t owns slots [t S, (t+1) S) = S/2 pair positions; pre[j * nthreads + t] = product of the denominators
// of its pair positions before j (1 for positions that hold no addition), tot[t] = the product of all of them.
template <class F, int S>
__global__ void __launch_bounds__(128)
k_ba_prefix(const Affine<F>* __restrict__ pts, const uint64_t* __restrict__ pairs, const uint32_t* __restrict__ d_n,
            uint32_t shift, F* pre, F* tot, uint32_t nthreads) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const size_t n = *d_n >> shift;
    F acc = F::one();
    size_t lo = (size_t)t * S;
    if (lo < n) {
        const uint64_t* pp = pairs + lo;
#pragma unroll 2
        for (int j = 0; j < S / 2; j++) {
            if (lo + 2 * j >= n) break;
            const uint64_t pa = pp[2 * j], pb = pp[2 * j + 1];
            pre[(size_t)j * nthreads + t] = acc;
            if (pb == BA_EMPTY) continue;                 // empty pair or odd leftover (pa valid, pb empty)
            acc = acc * ba_denominator(pts, (uint32_t)pa, (uint32_t)pb);
        }
    }
    tot[t] = acc;
}

template <class F, int S>
__global__ void __launch_bounds__(128)
k_ba_apply(const Affine<F>* __restrict__ pts, const uint64_t* __restrict__ pairs, const uint32_t* __restrict__ d_n,
           uint32_t shift, const F* __restrict__ pre, const F* __restrict__ inv_tot, Affine<F>* out_pts,
           uint64_t* out_pairs, uint32_t nthreads) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const size_t n = *d_n >> shift;
    size_t lo = (size_t)t * S;
    if (lo >= n) return;
    const uint64_t* pp = pairs + lo;
    F inv = inv_tot[t];          // 1 / (product of all denominators of this thread)
#pragma unroll 2
    for (int j = S / 2 - 1; j >= 0; j--) {
        if (lo + 2 * j >= n) continue;
        const uint64_t pa = pp[2 * j], pb = pp[2 * j + 1];
        const size_t slot = (lo >> 1) + j;
        if (pb == BA_EMPTY) {                             // nothing to add: empty pair, or an odd leftover passes through
            if (pa == BA_EMPTY) {
                out_pairs[slot] = BA_EMPTY;
            } else {
                out_pts[slot] = ba_load(pts, (uint32_t)pa);
                out_pairs[slot] = (pa & 0xffffffff00000000ull) | ((uint64_t)slot << 1);
            }
            continue;
        }
        Affine<F> p1 = ba_load(pts, (uint32_t)pa);
        Affine<F> p2 = ba_load(pts, (uint32_t)pb);
        F d;
        const int kind = ba_classify(p1, p2, d);
        F inv_d = inv * pre[(size_t)j * nthreads + t];
        inv = inv * d;
        out_pts[slot] = ba_add(p1, p2, inv_d, kind);
        out_pairs[slot] = (pa & 0xffffffff00000000ull) | ((uint64_t)slot << 1);
    }
}

// forward level of the product tree: groups of 64 consecutive values; pre[i] = product of the group's
// values before i (exclusive), group_tot[g] = product of the whole group
template <class F>
__global__ void __launch_bounds__(128)
k_ba_tree_fwd(const F* __restrict__ vals, size_t n, F* pre, F* group_tot) {
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = g * 64;
    if (lo >= n) return;
    size_t hi = lo + 64 < n ? lo + 64 : n;
    F acc = F::one();
    for (size_t i = lo; i < hi; i++) {
        pre[i] = acc;
        acc = acc * vals[i];
    }
    group_tot[g] = acc;
}
// backward level: group_inv[g] = 1 / (product of the group) -> vals[i] = 1 / vals[i]
template <class F>
__global__ void __launch_bounds__(128)
k_ba_tree_bwd(F* vals, size_t n, const F* __restrict__ pre, const F* __restrict__ group_inv) {
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = g * 64;
    if (lo >= n) return;
    size_t hi = lo + 64 < n ? lo + 64 : n;
    F inv = group_inv[g];
    for (size_t i = hi; i-- > lo;) {
        F v = vals[i];
        vals[i] = inv * pre[i];
        inv = inv * v;
    }
}
// top of the tree: one warp per value (lane 0 works: the GCD inversion's branches must not serialise)
template <class F>
__global__ void __launch_bounds__(128)
k_ba_invert_top(F* vals, size_t n) {
    size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g & 31) return;
    size_t i = g >> 5;
    if (i < n) vals[i] = vals[i].inverse_fast();
}

// vals[0..n) -> their inverses, in place (no value may be zero)
template <class F>
void ba_invert(hk_ctx* ctx, MsmWs& ws, cudaStream_t st, F* vals, size_t n) {
    std::vector<size_t> sizes{n};
    while (sizes.back() > 1024) sizes.push_back((sizes.back() + 63) / 64);
    size_t pre_total = 0, lvl_total = 0;
    for (size_t l = 0; l + 1 < sizes.size(); l++) pre_total += sizes[l], lvl_total += sizes[l + 1];
    F* pre = ws.ba_pre.as<F>(pre_total + 1);
    F* lvl = ws.ba_lvl.as<F>(lvl_total + 1);
    std::vector<F*> v{vals}, p;
    {
        F* pp = pre;
        F* ll = lvl;
        for (size_t l = 0; l + 1 < sizes.size(); l++) {
            p.push_back(pp);
            pp += sizes[l];
            v.push_back(ll);
            ll += sizes[l + 1];
        }
    }
    for (size_t l = 0; l + 1 < sizes.size(); l++) {
        k_ba_tree_fwd<F><<<ceil_div(sizes[l + 1], 128), 128, 0, st>>>(v[l], sizes[l], p[l], v[l + 1]);
        HK_LAUNCH_CHECK(ctx);
    }
    k_ba_invert_top<F><<<ceil_div(sizes.back() * 32, 128), 128, 0, st>>>(v.back(), sizes.back());
    HK_LAUNCH_CHECK(ctx);
    for (size_t l = sizes.size() - 1; l-- > 0;) {
        k_ba_tree_bwd<F><<<ceil_div(sizes[l + 1], 128), 128, 0, st>>>(v[l], sizes[l], p[l], v[l + 1]);
        HK_LAUNCH_CHECK(ctx);
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
    if (threadIdx.x != 0) return;
    const XYZZ<F>* wsum = wsum_all + (size_t)blockIdx.x * W;
    XYZZ<F> acc = wsum[W - 1];
    for (int w = W - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        acc.add(wsum[w]);
    }
    out[blockIdx.x] = Jacobian<F>::from_xyzz(acc);
}

template <class F>
__global__ void k_set_identity_jac(Jacobian<F>* out, size_t K) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < K) out[i] = {F::one(), F::one(), F::zero()};
}

template <class Curve, int GROUP>
void bases_precompute(hk_ctx* ctx, BaseSet& bs, double density) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
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
    bs.tbl = t;
    bs.tbl_c = pl.c;
    bs.tbl_W = pl.W;
}

template <class Curve, int GROUP>
void msm_run(hk_ctx* ctx, const MsmBases& bv, size_t offset, size_t n, const Scalar256* d_scalars_in, size_t K,
             bool scalars_mont, void* d_out_jac_v, size_t sc_stride, int lane) {
    using F = typename GroupFieldT<Curve, GROUP>::F;
    using Fr = typename Curve::Fr;
    Jacobian<F>* d_out = reinterpret_cast<Jacobian<F>*>(d_out_jac_v);
    if (lane < 0 || lane >= hk_ctx::LANES) lane = 0;
    cudaStream_t st = ctx->lane_stream(lane);      // control stream: sort, folds, reduction
    cudaStream_t sb = ctx->bulk_stream(lane);      // the accumulation kernel (== st on lane 0)
    MsmWs& ws = ctx->mws[lane];
    if (K == 0) return;
    if (n == 0) {
        k_set_identity_jac<F><<<ceil_div(K, 128), 128, 0, st>>>(d_out, K);
        HK_LAUNCH_CHECK(ctx);
        return;
    }
    if (sc_stride == 0) sc_stride = n;
    HK_REQUIRE(n < (size_t(1) << 31), HK_ERR_UNSUPPORTED, "msm: n must be < 2^31");
    const bool table = bv.tbl != nullptr;
    MsmPlan pl = table ? msm_plan_for(bv.c, Curve::FR_BITS) : msm_plan(n, Curve::FR_BITS, false);
    if (table) HK_REQUIRE(pl.W == bv.W, HK_ERR_INVALID, "msm: window table geometry mismatch");
    // bases pointer already advanced by `offset`; table rows are n_total apart
    const Affine<F>* d_bases = reinterpret_cast<const Affine<F>*>(table ? bv.tbl : bv.pts) + offset;
    const uint32_t per_window = table ? 0u : 1u;
    const uint32_t item_stride = table ? (uint32_t)bv.n_total : 0u;
    const uint32_t Wred = table ? 1u : (uint32_t)pl.W;        // bucket sets per scalar vector
    const uint32_t nsets = (uint32_t)K * Wred;
    const size_t nkeys = (size_t)nsets * pl.B;
    const size_t max_items = n * K * (size_t)pl.W;
    HK_REQUIRE(nkeys < (size_t(1) << 31), HK_ERR_UNSUPPORTED, "msm: too many buckets for one batch");
    HK_REQUIRE(max_items < (size_t(1) << 32), HK_ERR_UNSUPPORTED, "msm: batch * n * windows must be < 2^32");
    // level-0 segment length / occupancy variant (HK_ACCUM_S in {32,64,128}, HK_ACCUM_MINB in {2,3,4}, default 4:
    // with the call-based multiply 4 CTAs/SM at 128 registers beat 2 at 206 -- profiles/r01c;
    // G2 always uses MINB = 1)
    static const int env_minb = getenv("HK_ACCUM_MINB") ? atoi(getenv("HK_ACCUM_MINB")) : 4;
    // level-0 segment length: 32 (64 / 128 measured slower: profiles/r01c); a standalone MSM whose list cannot
    // fill the GPU with 32-item segments (latency-bound 32-addition chains) uses 8
    static const int env_s0 = getenv("HK_ACCUM_S") ? atoi(getenv("HK_ACCUM_S")) : 0;
    const int S0 = env_s0 == 8 || env_s0 == 32 ? env_s0 : ((lane == 0 && n * K * (size_t)pl.W <= (size_t(1) << 21)) ? 8 : 32);
    // fold segment length: short segments mean more levels (each halves the list) but a 4-addition serial chain
    // per level instead of 16; measured per 2^20 G1 MSM 16 / 8 / 4 -> 7.83 / 7.66 / 7.64 ms, 2^16: 1.93 / 1.74 / 1.69,
    // G2 2^18: 10.10 / 9.49 / 9.23 (HK_FOLD_S overrides)
    // Inside a proof batch (lane != 0) the tails are hidden under other lanes' accumulations and the extra
    // launches of short segments cost throughput (280 -> 264 proofs/s), so only standalone MSMs use 4.
    static const int env_s1 = getenv("HK_FOLD_S") ? atoi(getenv("HK_FOLD_S")) : 0;
    const int S1 = (env_s1 == 4 || env_s1 == 8 || env_s1 == 16 || env_s1 == 32) ? env_s1 : (lane == 0 ? 4 : 16);

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
    // batch-affine pre-reduction: OFF by default (hk_ctx_set_option HK_OPT_BA_ROUNDS / HK_OPT_BA_MIN_ITEMS, or the
    // HK_BA_ROUNDS / HK_BA_MIN environment defaults).  Measured on B200 (profiles/r01e_batch_affine.md): a round costs
    // 2.58 ms per 8.4 M additions against 2.92 ms for the XYZZ kernel (+13 %), but adds ~0.65 ms of inversion-tree
    // latency per round and leaves a short, latency-bound list to the accumulation -- a wash in the proof
    // pipeline (275 vs 274 proofs/s) and a loss for a standalone MSM (9.1 vs 8.0 ms at 2^20).
    constexpr int BA_S = 32;           // slots per thread in the batch-affine kernels
    const int env_ba_rounds = ctx->ba_rounds;
    const bool ba_on = env_ba_rounds > 0 &&
                       (ctx->ba_min_items == 0 || (max_items >= ctx->ba_min_items && max_items / nkeys >= 8));
    const int ba_rounds = ba_on ? (env_ba_rounds < 1 ? 1 : (env_ba_rounds > 5 ? 5 : env_ba_rounds)) : 0;
    // with the rounds on, every bucket's run is padded to a multiple of 2^R slots
    const uint32_t ba_mask = (1u << ba_rounds) - 1u;
    const size_t list_ub = ba_rounds ? ((max_items + (size_t)ba_mask * nkeys + ba_mask) & ~(size_t)ba_mask) : max_items;
    HK_REQUIRE(list_ub < (size_t(1) << 32), HK_ERR_UNSUPPORTED, "msm: padded item list must be < 2^32");
    const size_t acc_ub = list_ub >> ba_rounds;      // upper bound of the list length the accumulation kernel sees
    uint32_t* hist = ws.hist.as<uint32_t>(nkeys + 4);
    uint32_t* d_count = hist + nkeys;     // total item count lives right after the histogram
    uint64_t* pairs = ws.items.as<uint64_t>(list_ub + 2);
    XYZZ<F>* buckets = ws.buckets.as<XYZZ<F>>(nkeys);
    const uint32_t nthr0 = ceil_div(acc_ub, S0);
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
    XYZZ<F>* partial = ws.chunks.as<XYZZ<F>>(part_total);
    uint32_t* partial_keys = ws.misc.as<uint32_t>(part_total);
    // reduction geometry: short serial chunks, then a multi-block tree
    static const int env_g = getenv("HK_REDUCE_G") ? atoi(getenv("HK_REDUCE_G")) : 0;
    // standalone (lane 0): short chunks shorten the serial tail (8 / 4 / 2 / 1 at B = 32768 -> 7.97 / 7.88 / 7.86 / 7.95 ms
    // per 2^20 MSM); inside a proof batch the 4x larger grids of a latency-bound kernel take SM slots from the
    // accumulations (283 -> 269 proofs/s), so the lanes keep long chunks
    uint32_t G = lane == 0 ? (pl.B >= 64 ? 2 : 1) : (pl.B >= 32768 ? 8 : (pl.B >= 1024 ? 4 : (pl.B >= 64 ? 2 : 1)));
    if (env_g > 0 && (uint32_t)env_g <= pl.B) G = (uint32_t)env_g;
    uint32_t T = pl.B / G;
    constexpr int TS_TPB = 64;
    static_assert(TS_TPB * sizeof(XYZZ<F>) <= 48 * 1024, "tree-sum tile must fit the default dynamic smem limit");
    std::vector<uint32_t> tree;        // tree levels: T -> ceil(T / 128) -> ... -> 1 (per bucket set)
    for (uint32_t m = T; m > 1; m = ceil_div(m, 2 * TS_TPB)) tree.push_back(m);
    size_t tree_total = (size_t)T * nsets;
    for (uint32_t m : tree) tree_total += (size_t)ceil_div(m, 2 * TS_TPB) * nsets;
    XYZZ<F>* chunk_out = ws.wsum.as<XYZZ<F>>(tree_total + nsets);

    HK_CUDA(cudaMemsetAsync(hist, 0, (nkeys + 4) * sizeof(uint32_t), st));
    HK_CUDA(cudaMemsetAsync(buckets, 0, nkeys * sizeof(XYZZ<F>), st));
    const uint8_t* skip = bv.skip ? bv.skip + offset : nullptr;
    k_msm_hist<<<ceil_div(n * K, 256), 256, 0, st>>>(d_sc, n, K, sc_stride, pl.c, pl.W, pl.B, Wred, per_window, skip, hist);
    HK_LAUNCH_CHECK(ctx);
    if (ba_rounds) {
        k_ba_pad_counts<<<ceil_div(nkeys, 256), 256, 0, st>>>(hist, nkeys, ba_mask);
        HK_LAUNCH_CHECK(ctx);
    }
    exclusive_scan_u32(ctx, ws, st, hist, nkeys, d_count);
    if (ba_rounds) {
        k_ba_fill_empty<<<ctx->sm_count * 8, 256, 0, st>>>(pairs, d_count);
        HK_LAUNCH_CHECK(ctx);
    }
    k_msm_scatter<<<ceil_div(n * K, 256), 256, 0, st>>>(d_sc, n, K, sc_stride, pl.c, pl.W, pl.B, Wred, per_window,
                                                        item_stride, skip, hist, pairs);
    HK_LAUNCH_CHECK(ctx);
    if (sb != st) {
        HK_CUDA(cudaEventRecord(ctx->ev_sorted[lane], st));
        HK_CUDA(cudaStreamWaitEvent(sb, ctx->ev_sorted[lane], 0));
    }
    const Affine<F>* acc_bases = d_bases;
    const uint64_t* acc_pairs = pairs;
    for (int r = 0; r < ba_rounds; r++) {
        const size_t n_ub = list_ub >> r;
        const uint32_t nthr = ceil_div(n_ub, BA_S);
        F* tot = ws.ba_tot.as<F>(nthr);
        F* pre = ws.ba_head.as<F>((size_t)nthr * (BA_S / 2));
        Affine<F>* out_pts = ws.ba_pts[r & 1].as<Affine<F>>((n_ub >> 1) + 1);
        uint64_t* out_pairs = ws.ba_pairs[r & 1].as<uint64_t>((n_ub >> 1) + 2);
        int pb = prof_begin(ctx, GROUP == 1 ? HK_PROF_MSM_BA_G1 : HK_PROF_MSM_BA_G2, sb);
        k_ba_prefix<F, BA_S><<<ceil_div(nthr, 128), 128, 0, sb>>>(acc_bases, acc_pairs, d_count, (uint32_t)r, pre, tot, nthr);
        HK_LAUNCH_CHECK(ctx);
        ba_invert<F>(ctx, ws, sb, tot, nthr);
        k_ba_apply<F, BA_S><<<ceil_div(nthr, 128), 128, 0, sb>>>(acc_bases, acc_pairs, d_count, (uint32_t)r, pre, tot, out_pts,
                                                                out_pairs, nthr);
        HK_LAUNCH_CHECK(ctx);
        prof_end(ctx, pb, 0, d_count, sb);
        acc_bases = out_pts;
        acc_pairs = out_pairs;
    }
    int pi = prof_begin(ctx, GROUP == 1 ? HK_PROF_MSM_ACCUM_G1 : HK_PROF_MSM_ACCUM_G2, sb);
    {
        const dim3 grid(ceil_div(nthr0, 128));
#define HK_ACC(SS, MB)                                                                                               \
    do {                                                                                                            \
        if (ba_rounds)                                                                                              \
            k_msm_accum<F, SS, MB, true><<<grid, 128, 0, sb>>>(acc_bases, acc_pairs, d_count, (uint32_t)ba_rounds, buckets, \
                                                               partial, partial_keys, nthr0);                       \
        else                                                                                                        \
            k_msm_accum<F, SS, MB, false><<<grid, 128, 0, sb>>>(acc_bases, acc_pairs, d_count, 0u, buckets, partial, \
                                                                partial_keys, nthr0);                               \
    } while (0)
        if constexpr (GROUP == 2) {
            if (S0 == 8) HK_ACC(8, 1); else HK_ACC(32, 1);
        } else {
            const int mb = env_minb;
            if (S0 == 8) HK_ACC(8, 4);
            else if (mb == 3) HK_ACC(32, 3); else if (mb == 2) HK_ACC(32, 2); else HK_ACC(32, 4);
        }
#undef HK_ACC
    }
    HK_LAUNCH_CHECK(ctx);
    prof_end(ctx, pi, 0, d_count, sb);
    if (sb != st) {
        HK_CUDA(cudaEventRecord(ctx->ev_accum[lane], sb));
        HK_CUDA(cudaStreamWaitEvent(st, ctx->ev_accum[lane], 0));
    }
    size_t off = 0;
    for (size_t l = 0; l < lvl_in.size(); l++) {
        uint32_t n_in = lvl_in[l];
        uint32_t thr = ceil_div(n_in, S1);
        int final_level = (l + 1 == lvl_in.size());
        XYZZ<F>* in = partial + off;
        uint32_t* in_keys = partial_keys + off;
        XYZZ<F>* out = partial + off + n_in;
        uint32_t* out_keys = partial_keys + off + n_in;
        if (S1 == 8) k_msm_fold<F, 8><<<ceil_div(thr, 128), 128, 0, st>>>(in, in_keys, n_in, buckets, out, out_keys, thr, final_level);
        else if (S1 == 4) k_msm_fold<F, 4><<<ceil_div(thr, 128), 128, 0, st>>>(in, in_keys, n_in, buckets, out, out_keys, thr, final_level);
        else if (S1 == 32) k_msm_fold<F, 32><<<ceil_div(thr, 128), 128, 0, st>>>(in, in_keys, n_in, buckets, out, out_keys, thr, final_level);
        else k_msm_fold<F, 16><<<ceil_div(thr, 128), 128, 0, st>>>(in, in_keys, n_in, buckets, out, out_keys, thr, final_level);
        HK_LAUNCH_CHECK(ctx);
        off += n_in;
    }
    k_msm_reduce_chunks<F><<<ceil_div((size_t)T * nsets, 128), 128, 0, st>>>(buckets, pl.B, G, T, nsets, chunk_out);
    HK_LAUNCH_CHECK(ctx);
    XYZZ<F>* cur = chunk_out;
    uint32_t cur_stride = T;
    XYZZ<F>* nxt = chunk_out + (size_t)T * nsets;
    for (uint32_t m : tree) {
        uint32_t nb = ceil_div(m, 2 * TS_TPB);
        k_msm_tree_sum<F, TS_TPB><<<dim3(nb, nsets), TS_TPB, TS_TPB * sizeof(XYZZ<F>), st>>>(cur, m, cur_stride, nxt, nb);
        HK_LAUNCH_CHECK(ctx);
        cur = nxt;
        cur_stride = nb;
        nxt += (size_t)nb * nsets;
    }
    // cur: one point per bucket set, K * Wred of them, vector-major
    k_msm_combine<F><<<(unsigned)K, 32, 0, st>>>(cur, (int)Wred, pl.c, d_out);
    HK_LAUNCH_CHECK(ctx);
}

}  // namespace hk
