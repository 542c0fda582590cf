// CPU fuzz of hk::fold_runs (hekaton_system_b200/csrc/fold_runs.cuh): random partial lists with the structure the
// accumulation / fold kernels produce, checked against a brute-force grouping of the pieces.  Test infrastructure.
//   * lists built from a sorted item stream cut into equal segments (what k_msm_accum emits), with run lengths from
//     "all distinct" to "one bucket holds everything";
//   * if the brute force finds a run of >= 3 pieces (or a hole was injected) some thread must report `bad`;
//   * otherwise the runs reported by all threads must be exactly the brute-force groups, each once.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <utility>
#include <vector>

#include "fold_runs.cuh"

using hk::KEY_NONE;

struct Piece {
    uint32_t slot, key;
};

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937_64 rng(0x48454b41);
    long n_bad = 0, n_ok = 0, n_runs = 0, n_merges = 0;
    for (int round = 0; round < rounds; round++) {
        const uint32_t n = 1 + rng() % 40;                    // producer threads
        const uint32_t S = 1 + rng() % 6;                     // items per thread
        const uint32_t count = rng() % (n * S + 1);           // items actually present (threads past the end are empty)
        const uint32_t spread = 1 + rng() % (rng() % 2 ? 4 : 64);   // small: long runs, large: mostly distinct keys
        std::vector<uint32_t> items(count);
        uint32_t k = rng() % 5;
        for (auto& v : items) {
            if (rng() % spread) k += rng() % 3 == 0 ? 1 + rng() % 3 : 1;
            v = k;
        }
        std::vector<uint32_t> keys(2 * n, KEY_NONE);
        for (uint32_t t = 0; t < n; t++) {
            const uint32_t lo = t * S, hi = lo + S < count ? lo + S : count;
            if (lo >= count) break;
            keys[2 * t] = items[lo];
            if (items[hi - 1] != items[lo]) keys[2 * t + 1] = items[hi - 1];
        }
        bool hole = false;
        if (rng() % 8 == 0 && n > 2) {                        // inject a hole in front of a non-empty thread
            const uint32_t t = rng() % (n - 1);
            if (keys[2 * t + 2] != KEY_NONE) {
                keys[2 * t] = keys[2 * t + 1] = KEY_NONE;
                hole = true;
            }
        }
        // brute force: maximal groups of adjacent pieces with equal keys
        std::vector<Piece> pieces;
        for (uint32_t s = 0; s < 2 * n; s++)
            if (keys[s] != KEY_NONE) pieces.push_back({s, keys[s]});
        std::vector<std::vector<uint32_t>> groups;
        for (size_t p = 0; p < pieces.size(); p++) {
            if (p == 0 || pieces[p].key != pieces[p - 1].key) groups.emplace_back();
            groups.back().push_back(pieces[p].slot);
        }
        bool long_run = false;
        for (auto& g : groups) long_run |= g.size() >= 3;
        bool any_bad = false;
        std::map<uint32_t, std::vector<uint32_t>> got;         // start slot -> slots of the run
        for (uint32_t i = 0; i < n; i++) {
            const hk::FoldRuns r = hk::fold_runs(keys.data(), i, n);
            any_bad |= r.bad;
            if (r.a) {
                auto& v = got[2 * i];
                v.push_back(2 * i);
                if (r.a_merge) v.push_back(2 * i + 2);
                if (r.ka != keys[2 * i]) return printf("round %d: wrong key for run A of thread %u\n", round, i), 1;
            }
            if (r.b) {
                auto& v = got[2 * i + 1];
                v.push_back(2 * i + 1);
                if (r.b_merge) v.push_back(2 * i + 2);
                if (r.kb != keys[2 * i + 1]) return printf("round %d: wrong key for run B of thread %u\n", round, i), 1;
            }
        }
        if (hole || long_run) {
            if (!any_bad) return printf("round %d: %s not reported\n", round, hole ? "hole" : "run of >= 3 pieces"), 1;
            n_bad++;
            continue;
        }
        if (any_bad) return printf("round %d: `bad` on a list whose runs all have <= 2 pieces\n", round), 1;
        if (got.size() != groups.size()) return printf("round %d: %zu runs reported, %zu expected\n", round, got.size(), groups.size()), 1;
        for (auto& g : groups) {
            auto it = got.find(g[0]);
            if (it == got.end() || it->second != g) return printf("round %d: run starting at slot %u differs\n", round, g[0]), 1;
            n_merges += g.size() == 2;
        }
        n_runs += (long)groups.size();
        n_ok++;
    }
    printf("ok: %ld lists resolved (%ld runs, %ld merges), %ld lists rejected\n", n_ok, n_runs, n_merges, n_bad);
    return 0;
}
