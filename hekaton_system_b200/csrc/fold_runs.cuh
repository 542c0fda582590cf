// Classification of the runs that start in one producer thread's slots of a partial list (msm_impl.cuh,
// k_msm_fold_check / k_msm_fold_resolve_quad).  Plain C++: compiled by nvcc for the device and by g++ for the
// CPU fuzz test (tests/test_fold_runs_host.py).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HK_FOLD_HD __host__ __device__ __forceinline__
#else
#define HK_FOLD_HD inline
#endif

namespace hk {

static constexpr uint32_t KEY_NONE = 0xffffffffu;      // empty slot of a partial list

// Thread i of the producing level left the sum of the first run of its range in slot 2i ("F_i", key keys[2i]) and of
// its last run in slot 2i+1 ("L_i"; KEY_NONE when the whole range was one run).  Keys never decrease along the list
// and empty threads only follow the last non-empty one.
struct FoldRuns {
    bool a, a_merge, b, b_merge, bad;      // run A starts at F_i (pieces F_i [, F_{i+1}]); run B at L_i ([, F_{i+1}])
    uint32_t ka, kb;
};

// The (at most two) runs of equal keys that START in thread i's slots, each with at most one more piece (F_{i+1});
// `bad` if such a run has a third piece (thread i+1 holds a single piece of the same key and F_{i+2} continues it) or
// if thread i is a hole in front of a non-empty thread.  Every run of the list starts in exactly one thread, so the
// union over i covers every piece exactly once when no thread reports `bad`.
HK_FOLD_HD FoldRuns fold_runs(const uint32_t* __restrict__ keys, uint32_t i, uint32_t n) {
    FoldRuns r{false, false, false, false, false, KEY_NONE, KEY_NONE};
    const uint32_t k0 = keys[2 * i], k1 = keys[2 * i + 1];
    const uint32_t nk0 = i + 1 < n ? keys[2 * i + 2] : KEY_NONE, nk1 = i + 1 < n ? keys[2 * i + 3] : KEY_NONE;
    if (k0 == KEY_NONE) {                  // empty thread: legal only past the end of the list
        r.bad = (k1 != KEY_NONE) || (nk0 != KEY_NONE);
        return r;
    }
    uint32_t prev = KEY_NONE;              // key of the piece in front of F_i
    if (i > 0) {
        prev = keys[2 * i - 1];
        if (prev == KEY_NONE) prev = keys[2 * i - 2];
    }
    const uint32_t nnk0 = i + 2 < n ? keys[2 * i + 4] : KEY_NONE;
    const bool next_single = nk0 != KEY_NONE && nk1 == KEY_NONE;      // thread i+1 holds one piece only
    r.ka = k0;
    r.kb = k1;
    if (k1 != KEY_NONE) {
        r.a = prev != k0;                  // F_i alone (L_i has another key)
        r.b = true;
        r.b_merge = nk0 == k1;
        r.bad = r.b_merge && next_single && nnk0 == k1;
    } else if (prev != k0) {
        r.a = true;
        r.a_merge = nk0 == k0;
        r.bad = r.a_merge && next_single && nnk0 == k0;
    }
    return r;
}

}  // namespace hk
