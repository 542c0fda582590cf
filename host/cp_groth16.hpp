// C++17 mirror of the reference's cp-groth16 operator interface for the hot path, over the C ABI
// (include/hekaton_b200.h).  Header-only; links against libhekaton_b200.so only.
//
//   reference (Rust, /root/reference/cp-groth16/src)           here
//   ---------------------------------------------------------  -----------------------------------------
//   MultiStageConstraintSystem   constraint_synthesizer.rs:14   cpg16::MultiStageConstraintSystem
//     initialize_stage / finalize_stage / synthesize_with :53-77   same names
//     current_stage_witness_assignment :96-99, full_assignment :102-106
//   trait MultiStageConstraintSynthesizer  :119-134              cpg16::MultiStageConstraintSynthesizer
//   ProvingKey<E>                data_structures.rs:66-83        cpg16::ProvingKey (device-resident)
//   Proof<E>{a,b,c,ds}           data_structures.rs:7-16         cpg16::Proof
//   CommitmentBuilder::{new,commit,prove}  committer.rs:38-122   cpg16::CommitmentBuilder<C>
//
// Constraint synthesis stays on the host exactly as in the reference; what `generate_constraints`
// leaves behind (assignment + per-stage witness ranges) is what crosses into the device path.
// Randomness is drawn from the caller's Rng in the reference's order (kappa in commit:
// committer.rs:85; r then s in prove: prover.rs:28-29).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/hekaton_b200.h"

namespace cpg16 {

using Fr = std::array<uint64_t, 4>;   // Montgomery form, little-endian limbs (ark-ff layout)

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check(hk_ctx* ctx, int32_t rc, const char* what) {
    if (rc != HK_OK) throw Error(rc, std::string(what) + ": " + hk_last_error_string(ctx));
}

// Source of field randomness (the reference passes `&mut impl Rng` and calls Fr::rand).
struct Rng {
    virtual ~Rng() = default;
    virtual Fr fr_rand() = 0;
};

// SynthesisError::AssignmentMissing etc. are host-side concerns; only what the path needs is kept.
class MultiStageConstraintSystem {
  public:
    std::vector<Fr> instance_assignment;   // includes the constant one at index 0
    std::vector<Fr> witness_assignment;
    std::vector<std::pair<size_t, size_t>> variable_range_for_stage;

    void initialize_stage() {
        size_t start = witness_assignment.size();
        variable_range_for_stage.emplace_back(start, start);
    }
    void finalize_stage() { variable_range_for_stage.back().second = witness_assignment.size(); }
    void synthesize_with(const std::function<void(MultiStageConstraintSystem&)>& constraints) {
        initialize_stage();
        constraints(*this);
        finalize_stage();
    }
    size_t num_instance_variables() const { return instance_assignment.size(); }
    size_t num_witness_variables() const { return witness_assignment.size(); }
    std::vector<Fr> current_stage_witness_assignment() const {
        auto r = variable_range_for_stage.back();
        return std::vector<Fr>(witness_assignment.begin() + r.first, witness_assignment.begin() + r.second);
    }
    std::vector<Fr> full_assignment() const {
        std::vector<Fr> z(instance_assignment);
        z.insert(z.end(), witness_assignment.begin(), witness_assignment.end());
        return z;
    }
    // ark-relations' `new_input` / `new_witness` as far as assignments are concerned
    void new_input(const Fr& v) { instance_assignment.push_back(v); }
    void new_witness(const Fr& v) { witness_assignment.push_back(v); }
};

struct MultiStageConstraintSynthesizer {
    virtual ~MultiStageConstraintSynthesizer() = default;
    virtual size_t total_num_stages() const = 0;
    size_t last_stage() const { return total_num_stages() - 1; }
    virtual void generate_constraints(size_t stage, MultiStageConstraintSystem& cs) = 0;
};

struct G1Affine {
    std::vector<uint64_t> xy;   // x || y, Montgomery
    bool infinity = false;
    bool operator==(const G1Affine& o) const { return infinity == o.infinity && (infinity || xy == o.xy); }
};
struct G2Affine {
    std::vector<uint64_t> xy;   // x.c0 x.c1 y.c0 y.c1
    bool infinity = false;
    bool operator==(const G2Affine& o) const { return infinity == o.infinity && (infinity || xy == o.xy); }
};
using Comm = G1Affine;
using CommRandomness = Fr;

struct Proof {
    G1Affine a;
    G2Affine b;
    G1Affine c;
    std::vector<G1Affine> ds;
};

// Device-resident proving key (hk_pk_create).  Contexts forked from `ctx` can use `handle` too.
struct ProvingKey {
    hk_ctx* ctx = nullptr;
    uint64_t handle = 0;
    uint32_t num_stages = 0;
    uint32_t fq_limbs = 6;
    size_t num_variables = 0;
    size_t deltas_g_len() const { return num_stages; }
};

template <class C>
class CommitmentBuilder {
  public:
    MultiStageConstraintSystem cs;
    C circuit;

    // `ctx` is the context this builder launches on: pk.ctx or a fork of it (one per host thread).
    CommitmentBuilder(C circuit_, const ProvingKey& pk_, hk_ctx* ctx_ = nullptr)
        : circuit(std::move(circuit_)), pk(pk_), ctx(ctx_ ? ctx_ : pk_.ctx) {}

    std::pair<Comm, CommRandomness> commit(Rng& rng) {
        circuit.generate_constraints(cur_stage, cs);
        if (cur_stage >= pk.num_stages) throw Error(HK_ERR_INVALID, "no more values left in committing key");
        CommRandomness randomness = rng.fr_rand();
        std::vector<Fr> z = cs.full_assignment();
        Comm com;
        com.xy.resize(2 * pk.fq_limbs);
        uint8_t inf = 0;
        check(ctx, hk_commit(ctx, pk.handle, (uint32_t)cur_stage, z[0].data(), randomness.data(), HK_HOST, com.xy.data(), &inf),
              "hk_commit");
        com.infinity = inf != 0;
        cur_stage += 1;
        return {com, randomness};
    }

    Proof prove(const std::vector<Comm>& comms, const std::vector<CommRandomness>& comm_rands, Rng& rng) {
        // prove_last_stage_with_zk: r, s drawn first (prover.rs:28-29), then the last stage is synthesised
        Fr r = rng.fr_rand();
        Fr s = rng.fr_rand();
        circuit.generate_constraints(circuit.last_stage(), cs);
        if (pk.deltas_g_len() != comm_rands.size() + 1)
            throw Error(HK_ERR_LENGTH, "assertion failed: pk.deltas_g.len() == comm_rands.len() + 1");
        std::vector<Fr> z = cs.full_assignment();
        if (z.size() != pk.num_variables) throw Error(HK_ERR_LENGTH, "assignment length != number of key variables");
        Proof p;
        p.a.xy.resize(2 * pk.fq_limbs);
        p.b.xy.resize(4 * pk.fq_limbs);
        p.c.xy.resize(2 * pk.fq_limbs);
        uint8_t inf[3] = {0, 0, 0};
        check(ctx,
              hk_prove(ctx, pk.handle, z[0].data(), r.data(), s.data(), comm_rands.empty() ? nullptr : comm_rands[0].data(),
                       HK_HOST, p.a.xy.data(), p.b.xy.data(), p.c.xy.data(), inf),
              "hk_prove");
        p.a.infinity = inf[0];
        p.b.infinity = inf[1];
        p.c.infinity = inf[2];
        p.ds = comms;
        return p;
    }

  private:
    size_t cur_stage = 0;
    ProvingKey pk;
    hk_ctx* ctx;
};

}  // namespace cpg16
