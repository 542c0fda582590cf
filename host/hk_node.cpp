// `node work`-style driver over the C++ mirror: loads a workload file (matrices, key scalars,
// per-subcircuit assignments + randomness + expected outputs, written by
// hekaton_system_b200/workload.py::dump_workload), builds the device key, then proves every
// subcircuit with T host threads (one forked context each, like the per-proof thread pools of
// /root/reference/mpi-snark/src/bin/node.rs:745-795) and checks every commitment and proof
// byte-for-byte against the expected values.
//
//   hk_node <workload.bin> [threads]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <thread>

#include "cp_groth16.hpp"

using namespace cpg16;

struct Reader {
    std::ifstream f;
    explicit Reader(const char* p) : f(p, std::ios::binary) {
        if (!f) throw std::runtime_error(std::string("cannot open ") + p);
    }
    uint64_t u64() {
        uint64_t v;
        f.read((char*)&v, 8);
        return v;
    }
    template <class T>
    std::vector<T> vec(size_t n) {
        std::vector<T> v(n);
        if (n) f.read((char*)v.data(), n * sizeof(T));
        if (!f) throw std::runtime_error("short read");
        return v;
    }
};

// A circuit that replays a stored assignment stage by stage -- the role
// PartitionedR1CSCircuit plays in the reference (distributed-prover/src/partitioned_r1cs_circuit.rs).
struct ReplayCircuit : MultiStageConstraintSynthesizer {
    const std::vector<Fr>* z = nullptr;       // full assignment: instance || witness
    size_t l = 0;
    std::vector<uint64_t> stage_start;
    size_t total_num_stages() const override { return stage_start.size() - 1; }
    void generate_constraints(size_t stage, MultiStageConstraintSystem& cs) override {
        if (stage == 0)
            for (size_t i = 0; i < l; i++) cs.new_input((*z)[i]);
        cs.synthesize_with([&](MultiStageConstraintSystem& c) {
            for (size_t i = stage_start[stage]; i < stage_start[stage + 1]; i++) c.new_witness((*z)[l + i]);
        });
    }
};

struct ReplayRng : Rng {
    std::vector<Fr> vals;
    size_t pos = 0;
    Fr fr_rand() override { return vals.at(pos++); }
};

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: hk_node <workload.bin> [threads]\n");
        return 2;
    }
    int threads = argc > 2 ? atoi(argv[2]) : 4;
    try {
        Reader rd(argv[1]);
        if (rd.u64() != 0x314c574b48ull) throw std::runtime_error("bad magic");   // "HKWL1"
        uint32_t curve = (uint32_t)rd.u64();
        size_t n = rd.u64(), l = rd.u64(), nw = rd.u64(), ns = rd.u64();
        auto stage_start = rd.vec<uint64_t>(ns + 1);
        const size_t L = curve == HK_CURVE_BLS12_381 ? 6 : 4;
        const size_t m = l + nw;
        hk_ctx* ctx = nullptr;
        int32_t rc = hk_ctx_create(0, curve, nullptr, &ctx);
        if (rc != HK_OK) {
            fprintf(stderr, "hk_ctx_create failed (%d): an sm_100 GPU is required, there is no CPU fallback\n", rc);
            return 1;
        }
        // matrices
        std::vector<uint32_t> ptr[3], col[3];
        std::vector<uint64_t> val[3];
        for (int k = 0; k < 3; k++) {
            size_t nnz = rd.u64();
            ptr[k] = rd.vec<uint32_t>(n + 1);
            col[k] = rd.vec<uint32_t>(nnz);
            val[k] = rd.vec<uint64_t>(4 * nnz);
        }
        uint64_t r1cs = 0;
        check(ctx, hk_r1cs_upload(ctx, n, l, nw, ptr[0].data(), col[0].data(), val[0].data(), ptr[1].data(), col[1].data(),
                                  val[1].data(), ptr[2].data(), col[2].data(), val[2].data(), &r1cs), "hk_r1cs_upload");
        // key: generators + exponent vectors -> device points (setup helper; generator.rs:128-196)
        auto gen1 = rd.vec<uint64_t>(2 * L), gen2 = rd.vec<uint64_t>(4 * L);
        auto scal = [&](size_t cnt) { return rd.vec<uint64_t>(4 * cnt); };
        uint32_t log_n = 0;
        while ((size_t(1) << log_n) < n + l) log_n++;
        const size_t N = size_t(1) << log_n;
        auto mk1 = [&](const std::vector<uint64_t>& s, size_t cnt) {
            uint64_t h = 0;
            check(ctx, hk_bases_from_scalars_g1(ctx, gen1.data(), s.data(), cnt, HK_SCALAR_BIGINT, HK_HOST, &h), "bases_from_scalars_g1");
            return h;
        };
        auto mk2 = [&](const std::vector<uint64_t>& s, size_t cnt) {
            uint64_t h = 0;
            check(ctx, hk_bases_from_scalars_g2(ctx, gen2.data(), s.data(), cnt, HK_SCALAR_BIGINT, HK_HOST, &h), "bases_from_scalars_g2");
            return h;
        };
        auto a_s = scal(m), b1_s = scal(m), b2_s = scal(m), h_s = scal(N - 1);
        hk_pk_desc d;
        memset(&d, 0, sizeof d);
        d.a_g = mk1(a_s, m);
        d.b_g = mk1(b1_s, m);
        d.b_h = mk2(b2_s, m);
        d.h_g = mk1(h_s, N - 1);
        d.num_stages = (uint32_t)ns;
        for (size_t s = 0; s < ns; s++) {
            size_t cnt = stage_start[s + 1] - stage_start[s];
            auto cs = scal(cnt);
            d.ck[s] = mk1(cs, cnt);
        }
        auto del_s = scal(ns);
        d.deltas_g = mk1(del_s, ns);
        auto loose1 = scal(2), loose2 = scal(2);     // [alpha, beta] in G1 ; [beta, delta_last] in G2
        uint64_t hl1 = mk1(loose1, 2), hl2 = mk2(loose2, 2);
        std::vector<uint64_t> p1(2 * 2 * L), p2(2 * 4 * L);
        check(ctx, hk_bases_download(ctx, hl1, 0, 2, p1.data()), "download");
        check(ctx, hk_bases_download(ctx, hl2, 0, 2, p2.data()), "download");
        d.alpha_g = p1.data();
        d.beta_g = p1.data() + 2 * L;
        d.beta_h = p2.data();
        d.last_delta_h = p2.data() + 4 * L;
        d.r1cs = r1cs;
        for (size_t s = 0; s <= ns; s++) d.stage_start[s] = (uint32_t)stage_start[s];
        ProvingKey pk;
        pk.ctx = ctx;
        pk.num_stages = (uint32_t)ns;
        pk.fq_limbs = (uint32_t)L;
        pk.num_variables = m;
        check(ctx, hk_pk_create(ctx, &d, &pk.handle), "hk_pk_create");

        // subcircuits
        size_t K = rd.u64();
        struct Sub {
            std::vector<Fr> z;
            ReplayRng rng;
            std::vector<uint64_t> expD, expA, expB, expC;
        };
        std::vector<Sub> subs(K);
        for (auto& s : subs) {
            auto zz = rd.vec<uint64_t>(4 * m);
            s.z.resize(m);
            memcpy(s.z.data(), zz.data(), zz.size() * 8);
            auto rv = rd.vec<uint64_t>(4 * 3);      // kappa, r, s in draw order
            s.rng.vals.resize(3);
            memcpy(s.rng.vals.data(), rv.data(), rv.size() * 8);
            s.expD = rd.vec<uint64_t>(2 * L);
            s.expA = rd.vec<uint64_t>(2 * L);
            s.expB = rd.vec<uint64_t>(4 * L);
            s.expC = rd.vec<uint64_t>(2 * L);
        }
        if (threads < 1) threads = 1;
        if ((size_t)threads > K) threads = (int)K;
        std::vector<hk_ctx*> forks(threads, ctx);
        for (int t = 1; t < threads; t++) check(ctx, hk_ctx_fork(ctx, nullptr, &forks[t]), "hk_ctx_fork");
        std::atomic<size_t> bad{0};
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) {
            pool.emplace_back([&, t] {
                for (size_t i = t; i < K; i += threads) {
                    Sub& s = subs[i];
                    ReplayCircuit circ;
                    circ.z = &s.z;
                    circ.l = l;
                    circ.stage_start = stage_start;
                    CommitmentBuilder<ReplayCircuit> cb(circ, pk, forks[t]);
                    auto cr = cb.commit(s.rng);                       // worker.rs:48-74 (stage 0)
                    Proof pr = cb.prove({cr.first}, {cr.second}, s.rng);   // worker.rs:76-85 (stage 1)
                    bool ok = cr.first.xy == s.expD && pr.a.xy == s.expA && pr.b.xy == s.expB && pr.c.xy == s.expC &&
                              pr.ds.size() == 1;
                    if (!ok) bad++;
                }
            });
        }
        for (auto& th : pool) th.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("{\"subcircuits\": %zu, \"threads\": %d, \"mismatches\": %zu, \"seconds\": %.4f, \"proofs_per_s\": %.2f, "
               "\"gpu_launches\": %llu}\n",
               K, threads, (size_t)bad, dt, K / dt, (unsigned long long)hk_launch_count(ctx));
        for (int t = 1; t < threads; t++) hk_ctx_destroy(forks[t]);
        hk_ctx_destroy(ctx);
        return bad ? 1 : 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "hk_node: %s\n", e.what());
        return 1;
    }
}
