/*
 * hekaton_b200.h -- C ABI of the B200-native (sm_100a) backend for Hekaton's per-subcircuit
 * CP-Groth16 hot path.  This is the drop-in boundary: the entry points are exactly what a
 * Rust `-sys` crate for cp-groth16 would bind (INTEGRATION.md shows the binding).
 *
 * What each entry point replaces in the reference (/root/reference):
 *   hk_msm_g1 / hk_msm_g2        ark-ec `VariableBaseMSM::{msm, msm_unchecked, msm_bigint}` as
 *                                called at cp-groth16/src/committer.rs:89,113 and
 *                                cp-groth16/src/prover.rs:117,129,168 (via calculate_coeff :158-171)
 *   hk_ntt                       ark-poly `EvaluationDomain::{fft_in_place, ifft_in_place}` on a
 *                                Radix2EvaluationDomain and its `get_coset(F::GENERATOR)`, reached
 *                                through `QAP::witness_map` at cp-groth16/src/prover.rs:123
 *   hk_qap_h, hk_r1cs_*,         ark-groth16 `LibsnarkReduction::witness_map_from_matrices`
 *   hk_qap_witness_map           (re-exported at cp-groth16/src/lib.rs:9-10; call prover.rs:123)
 *   hk_bases_upload_*            the proving-key vectors of cp-groth16/src/data_structures.rs:66-114
 *                                (a_g, b_g, b_h, h_g, ck.deltas_abc_g, deltas_g), uploaded once
 *   hk_pk_*, hk_commit, hk_prove the fused per-subcircuit path: CommitmentBuilder::commit
 *                                (committer.rs:55-98) and CommitmentBuilder::prove ->
 *                                prove_last_stage (committer.rs:100-122, prover.rs:53-156)
 *
 * Conventions
 *   - All functions return int32_t: 0 = HK_OK, negative = error (hk_last_error_string()).
 *     Nothing throws or aborts across the boundary.  There is NO CPU fallback: without a
 *     usable sm_100 device hk_ctx_create fails.
 *   - Field elements are little-endian arrays of u64 limbs in MONTGOMERY form with
 *     R = 2^(64*LIMBS) -- the in-memory layout of ark-ff 0.4 `Fp<MontBackend<_,N>,N>`:
 *     Fr: 4 limbs (both curves).  Fq: 6 limbs (BLS12-381) / 4 limbs (BN254).
 *     Fq2 = c0 || c1.  Scalars passed with HK_SCALAR_BIGINT are canonical integers
 *     (`BigInt<4>`, what `into_bigint()` yields; prover.rs:78-82).
 *   - Affine points: x || y (G2: x.c0 x.c1 y.c0 y.c1) plus an optional byte-per-point
 *     infinity array (arkworks `Affine{x,y,infinity}` is not repr(C): the shim copies
 *     fields, never transmutes).
 *   - MSM results are Jacobian (X, Y, Z), x = X/Z^2, y = Y/Z^3, identity Z = 0 -- the
 *     representation of ark-ec `short_weierstrass::Projective`.  Any representative of the
 *     group element may be returned; equality with the reference is after into_affine().
 *   - `memspace`: HK_HOST = pointers are host memory (copies are part of the call),
 *     HK_DEVICE = pointers are device memory on the context's GPU (no copies; field-element and
 *     point pointers must be 16-byte aligned -- HK_ERR_INVALID otherwise),
 *     HK_HOST_PINNED (hk_prove_batch / hk_prove inputs only) = host memory that is already
 *     page-locked: inputs are DMA'd straight from the caller's buffers, no staging copy.
 *     HK_HOST_PINNED_ASYNC (hk_commit_batch / hk_prove_batch) = inputs AND outputs are page-locked host
 *     buffers; the call only enqueues (upload on a copy stream into one of two staging buffers, kernels,
 *     download into the caller's output buffers) and returns.  Inputs must stay untouched and outputs are
 *     valid only after hk_ctx_sync.  Several calls may be in flight on one context: the upload of the next
 *     batch overlaps the kernels of the current one.
 *   - A context is bound to one GPU and one CUDA stream and is not thread-safe; use one
 *     context per host thread (the reference runs several proofs concurrently on separate
 *     threads, mpi-snark/src/bin/node.rs:773-793).  Calls with HK_HOST outputs
 *     synchronise the stream before returning; HK_DEVICE calls are asynchronous.
 */
#ifndef HEKATON_B200_H
#define HEKATON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hk_ctx hk_ctx;

enum { HK_CURVE_BLS12_381 = 0, HK_CURVE_BN254 = 1 };
enum { HK_HOST = 0, HK_DEVICE = 1, HK_HOST_PINNED = 2, HK_HOST_PINNED_ASYNC = 3 };
enum { HK_SCALAR_BIGINT = 0, HK_SCALAR_MONT = 1 };
enum { HK_NTT_FORWARD = 0, HK_NTT_INVERSE = 1 };
enum { HK_DOMAIN_PLAIN = 0, HK_DOMAIN_COSET_GENERATOR = 1 };

enum {
    HK_OK = 0,
    HK_ERR_INVALID = -1,     /* bad argument */
    HK_ERR_CUDA = -2,        /* CUDA runtime error (text in hk_last_error_string) */
    HK_ERR_NOMEM = -3,
    HK_ERR_LENGTH = -4,      /* length mismatch where the reference returns Err (msm) */
    HK_ERR_HANDLE = -5,      /* unknown handle */
    HK_ERR_UNSUPPORTED = -6, /* e.g. log_n above the field's two-adicity */
    HK_ERR_NO_DEVICE = -7    /* no sm_100 GPU: there is no CPU fallback */
};

/* Field ids for the element-wise test entry point. */
enum { HK_FIELD_FR = 0, HK_FIELD_FQ = 1, HK_FIELD_FQ2 = 2 };
enum { HK_FOP_MUL = 0, HK_FOP_ADD = 1, HK_FOP_SUB = 2, HK_FOP_INV = 3, HK_FOP_TO_MONT = 4, HK_FOP_FROM_MONT = 5,
       HK_FOP_INV_GCD = 6 /* binary-GCD inverse (the one affine normalisation uses) */,
       HK_FOP_SQR = 7 /* dedicated Montgomery squaring (falls back to a*a where the field has no spare bits) */,
       HK_FOP_MUL_KARA = 8 /* a*b through the Karatsuba-product / separate-reduction multiply */,
       HK_FOP_MULSUB2 = 9 /* a*b - b*(a+b) through the fused dual product (two multiplies where the field lacks it) */ };

/* ---- context ---------------------------------------------------------------------- */
/* `stream`: a cudaStream_t to launch on (e.g. torch's current stream), or NULL to create one. */
int32_t hk_ctx_create(int device, uint32_t curve, void* stream, hk_ctx** out);
/* A second context on the same GPU that SHARES the parent's uploaded resources (base sets,
 * matrices, proving keys, NTT tables -- same handles) but has its own stream and scratch
 * memory, so independent proofs overlap.  This is the analogue of the reference running several
 * proofs on separate OS threads against one `ProvingKey` (mpi-snark/src/bin/node.rs:745-795).
 * Handle-table access is mutex-protected; each context is still single-threaded. */
int32_t hk_ctx_fork(hk_ctx* parent, void* stream, hk_ctx** out);
int32_t hk_ctx_destroy(hk_ctx* ctx);
/* Tuning knobs of a context's MSM pipeline (forks inherit them at fork time; 0 restores the built-in default;
 * results never depend on them -- DESIGN.md 3.2):
 *   HK_OPT_ACCUM_WAVES    cap of the bucket-accumulation grid in waves of resident CTAs (default 0 = no cap: one
 *                         thread per minimum segment); with a cap the per-thread segment length follows from the
 *                         item count on the device
 *   HK_OPT_ACCUM_SEG_MIN  smallest segment length (default 32; 8 for short standalone lists)
 *   HK_OPT_FOLD_SEG       entries per thread in the partial-list fold levels: 4, 8, 16 or 32 (default 16 inside a
 *                         proof batch, 4 standalone)
 *   HK_OPT_REDUCE_CHUNK   buckets per thread in the bucket reduction (power of two)
 *   HK_OPT_SORT_GLOBAL_ATOMICS  1 = always sort the digits with the global-atomic kernels (default 0: base sets with a
 *                         window table use block-private shared-memory histograms)
 *   HK_OPT_G2_LANE_PAIR   G2 bucket accumulation with every Fq2 value spread over a lane pair (BLS12-381 only):
 *                         0 / 1 = on (default: 0-9 % faster at every size), 2 = off (one thread per point addition)
 *   HK_OPT_QUAD_MAX_THREADS  fold / reduction / tree-sum launches of at most this many logical threads run with four
 *                         lanes per point addition (default 32768 for a standalone MSM, never inside a proof
 *                         batch; 1 = never)
 *   HK_OPT_NO_FOLD_RESOLVE  1 = fold the partial list level by level to the end (default 0: where the quad-cooperative
 *                         kernels are in use and a thread's range has outgrown the bucket runs, one launch closes
 *                         every border at once -- after a check that no run has more than two pieces -- and the
 *                         remaining levels return immediately) */
enum { HK_OPT_ACCUM_WAVES = 1, HK_OPT_ACCUM_SEG_MIN = 2, HK_OPT_FOLD_SEG = 3, HK_OPT_REDUCE_CHUNK = 4,
       HK_OPT_SORT_GLOBAL_ATOMICS = 5, HK_OPT_G2_LANE_PAIR = 6, HK_OPT_QUAD_MAX_THREADS = 7, HK_OPT_NO_FOLD_RESOLVE = 8 };
int32_t hk_ctx_set_option(hk_ctx* ctx, uint32_t option, uint64_t value);
int32_t hk_ctx_sync(hk_ctx* ctx);
const char* hk_last_error_string(hk_ctx* ctx);
/* Limb counts for this context's curve (fq_limbs64 = 6 or 4; fr_limbs64 = 4). */
int32_t hk_curve_info(hk_ctx* ctx, uint32_t* fq_limbs64, uint32_t* fr_limbs64, uint32_t* fr_bits,
                      uint32_t* two_adicity);
/* Arithmetic variants compiled into the MSM kernels of this library (same results bit for bit):
 *   HK_BUILD_MUL2      a*b - c*d of the point additions as one fused dual product (BLS12-381 Fq; default on)
 *   HK_BUILD_MUL_KARA  Karatsuba-product multiply (-DHK_MUL_KARA; measured slower, off)
 *   HK_BUILD_FP2_DUAL  Fq2 multiply as two dual products (-DHK_FP2_DUAL; off) */
enum { HK_BUILD_MUL2 = 1, HK_BUILD_MUL_KARA = 2, HK_BUILD_FP2_DUAL = 4 };
uint32_t hk_build_flags(void);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t hk_launch_count(hk_ctx* ctx);

/* ---- bases (proving-key vectors; data_structures.rs:66-114) ------------------------- */
int32_t hk_bases_upload_g1(hk_ctx* ctx, const void* affine_xy_mont, const uint8_t* infinity, size_t n,
                           uint32_t memspace, uint64_t* handle);
int32_t hk_bases_upload_g2(hk_ctx* ctx, const void* affine_xy_mont, const uint8_t* infinity, size_t n,
                           uint32_t memspace, uint64_t* handle);
int32_t hk_bases_free(hk_ctx* ctx, uint64_t handle);
int32_t hk_bases_len(hk_ctx* ctx, uint64_t handle, size_t* n);
/* Build the window-shifted table tbl[w][i] = 2^(c w) * P_i for a base set (W x its memory).  MSMs
 * on a set with a table use ONE bucket set for all windows and need no final doublings.
 * hk_pk_create does this for every vector of the key.  density_permille (0 = 1000) is the expected
 * share of full-width scalars and steers the window size: witness vectors are mostly bits. */
int32_t hk_bases_precompute(hk_ctx* ctx, uint64_t handle, uint32_t density_permille);
/* Copy points back (affine, Montgomery; infinity as (0,0)); out holds n points. */
int32_t hk_bases_download(hk_ctx* ctx, uint64_t handle, size_t offset, size_t n, void* out);
/* Setup helper (generator.rs:128-196 `FixedBase::msm`, restricted to what a synthetic CRS
 * needs): out[i] = scalars[i] * G for an affine generator G; creates a base set. */
int32_t hk_bases_from_scalars_g1(hk_ctx* ctx, const void* gen_affine_mont, const uint64_t* scalars,
                                 size_t n, uint32_t scalar_repr, uint32_t memspace, uint64_t* handle);
int32_t hk_bases_from_scalars_g2(hk_ctx* ctx, const void* gen_affine_mont, const uint64_t* scalars,
                                 size_t n, uint32_t scalar_repr, uint32_t memspace, uint64_t* handle);

/* ---- MSM (VariableBaseMSM) ----------------------------------------------------------
 * result = sum_{i<n} scalars[i] * bases[base_offset + i].
 * `n` is the number of scalars supplied.  hk_msm_* follow `msm_unchecked`/`msm_bigint`:
 * if fewer than n bases remain after base_offset the sum silently covers the shorter
 * length (committer.rs:110-113 relies on this).  hk_msm_*_checked follows `msm`: it
 * returns HK_ERR_LENGTH unless exactly n_bases_expected == n (committer.rs:89).
 * n == 0 -> identity (prover.rs:117 unwrap_or(zero)).  out: 3 field elements (Jacobian). */
int32_t hk_msm_g1(hk_ctx* ctx, uint64_t bases, size_t base_offset, const uint64_t* scalars, size_t n,
                  uint32_t scalar_repr, uint32_t memspace, void* out_jacobian);
int32_t hk_msm_g2(hk_ctx* ctx, uint64_t bases, size_t base_offset, const uint64_t* scalars, size_t n,
                  uint32_t scalar_repr, uint32_t memspace, void* out_jacobian);
int32_t hk_msm_g1_checked(hk_ctx* ctx, uint64_t bases, size_t base_offset, size_t n_bases_expected,
                          const uint64_t* scalars, size_t n, uint32_t scalar_repr, uint32_t memspace,
                          void* out_jacobian);
/* K scalar vectors against the same base range (same-shape subcircuits share a key). */
int32_t hk_msm_g1_batch(hk_ctx* ctx, uint64_t bases, size_t base_offset, const uint64_t* scalars, size_t n,
                        size_t K, uint32_t scalar_repr, uint32_t memspace, void* out_jacobian);
int32_t hk_msm_g2_batch(hk_ctx* ctx, uint64_t bases, size_t base_offset, const uint64_t* scalars, size_t n,
                        size_t K, uint32_t scalar_repr, uint32_t memspace, void* out_jacobian);

/* Element-wise scalar multiplication out[i] = scalars[i] * points[i], normalised to affine: the reference's
 * `scalar_pairing` (distributed-prover/src/pairing_ops.rs:32-39), which the aggregator applies to the A / C / D
 * / prepared-input vectors of all proofs (aggregation.rs:236-242,288-310).  points: n affine points (Montgomery),
 * infinity: n flags or NULL, scalars: n x 4 u64, out_xy: n affine points, inf_out: n bytes (may be NULL).
 * Host or device pointers according to `memspace` (HK_HOST / HK_DEVICE). */
int32_t hk_scalar_mul_g1(hk_ctx* ctx, const void* points, const uint8_t* infinity, const uint64_t* scalars, size_t n,
                         uint32_t scalar_repr, uint32_t memspace, void* out_xy, uint8_t* inf_out);
int32_t hk_scalar_mul_g2(hk_ctx* ctx, const void* points, const uint8_t* infinity, const uint64_t* scalars, size_t n,
                         uint32_t scalar_repr, uint32_t memspace, void* out_xy, uint8_t* inf_out);

/* Jacobian -> affine (CurveGroup::into_affine; prover.rs:152-154).  in: n*3 elements,
 * out: n*2 elements (Montgomery), inf_out: n bytes (may be NULL). Host pointers. */
int32_t hk_g1_into_affine(hk_ctx* ctx, const void* jac, size_t n, void* out_xy, uint8_t* inf_out);
int32_t hk_g2_into_affine(hk_ctx* ctx, const void* jac, size_t n, void* out_xy, uint8_t* inf_out);

/* ---- NTT (Radix2EvaluationDomain) ---------------------------------------------------
 * In-place, natural order in and out, size 2^log_n, Montgomery Fr.
 * dir = HK_NTT_FORWARD:  fft_in_place  (on a coset: x_i *= g^i first)
 * dir = HK_NTT_INVERSE:  ifft_in_place (includes 1/N; on a coset: then x_i *= g^-i)
 * domain = HK_DOMAIN_COSET_GENERATOR uses offset F::GENERATOR (7 / 5). */
int32_t hk_ntt(hk_ctx* ctx, uint64_t* inout, uint32_t log_n, uint32_t dir, uint32_t domain, uint32_t memspace);
int32_t hk_ntt_batch(hk_ctx* ctx, uint64_t* inout, uint32_t log_n, size_t K, uint32_t dir, uint32_t domain,
                     uint32_t memspace);

/* ---- R1CS -> QAP witness map (LibsnarkReduction) ------------------------------------
 * hk_qap_h: a, b, c are the length-2^log_n evaluation vectors (zero padded;
 * a[n..n+l] = instance assignment); h receives all 2^log_n coefficients. */
int32_t hk_qap_h(hk_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                 uint32_t memspace, uint64_t* h);
/* Resident constraint matrices (CSR over columns [instance || witness]); coefficients
 * Montgomery Fr.  row_ptr has num_constraints+1 entries per matrix. */
int32_t hk_r1cs_upload(hk_ctx* ctx, size_t num_constraints, size_t num_instance, size_t num_witness,
                       const uint32_t* a_row_ptr, const uint32_t* a_col, const uint64_t* a_val,
                       const uint32_t* b_row_ptr, const uint32_t* b_col, const uint64_t* b_val,
                       const uint32_t* c_row_ptr, const uint32_t* c_col, const uint64_t* c_val,
                       uint64_t* handle);
int32_t hk_r1cs_free(hk_ctx* ctx, uint64_t handle);
/* witness_map(cs): z = full assignment (instance || witness), Montgomery Fr, length
 * num_instance + num_witness; h receives 2^log_n coefficients, log_n = ceil_log2(n + l). */
int32_t hk_qap_witness_map(hk_ctx* ctx, uint64_t r1cs, const uint64_t* z, uint32_t memspace, uint64_t* h);

/* ---- fused per-subcircuit path ------------------------------------------------------ */
typedef struct hk_pk_desc {
    uint64_t a_g, b_g, h_g;   /* G1 base-set handles (data_structures.rs:71-77) */
    uint64_t b_h;             /* G2 base-set handle */
    uint64_t ck[8];           /* ck.deltas_abc_g[stage], G1 handles */
    uint32_t num_stages;
    uint64_t deltas_g;        /* G1 handle, num_stages points */
    const void* alpha_g;      /* vk.alpha_g  (G1 affine, host) */
    const void* beta_g;       /* G1 affine, host */
    const void* beta_h;       /* vk.beta_h   (G2 affine, host) */
    const void* last_delta_h; /* vk.deltas_h.last() (G2 affine, host) */
    uint64_t r1cs;            /* hk_r1cs_upload handle */
    uint32_t stage_start[9];  /* witness index ranges: stage i = [stage_start[i], stage_start[i+1]) */
    uint32_t witness_density_permille; /* expected share of full-width assignment scalars; 0 = default (150) */
} hk_pk_desc;

int32_t hk_pk_create(hk_ctx* ctx, const hk_pk_desc* desc, uint64_t* pk_handle);
int32_t hk_pk_free(hk_ctx* ctx, uint64_t pk_handle);
/* committer.rs:85-91: com = MSM(ck[stage], w_stage) + kappa * last_delta_g  -> affine.
 * z: full assignment so far (Montgomery Fr; at least instance + stage witnesses).
 * kappa: Montgomery Fr.  out: G1 affine (x||y Montgomery), inf flag. */
int32_t hk_commit(hk_ctx* ctx, uint64_t pk, uint32_t stage, const uint64_t* z, const uint64_t* kappa,
                  uint32_t memspace, void* com_affine, uint8_t* com_inf);
/* K commitments of one stage against the same key in one pipeline; assignment k starts at
 * z + k * z_stride * 4 (z_stride in field elements, >= l + stage_start[stage+1]). */
int32_t hk_commit_batch(hk_ctx* ctx, uint64_t pk, uint32_t stage, size_t K, const uint64_t* z, size_t z_stride,
                        const uint64_t* kappa, uint32_t memspace, void* com_affine, uint8_t* com_inf);
/* prover.rs:53-156 + committer.rs:106-114.  z: full assignment (Montgomery Fr),
 * r, s: Montgomery Fr, kappas: num_stages-1 Montgomery Fr.
 * out: A (G1 affine), B (G2 affine), C (G1 affine), each x||y Montgomery; inf[3]. */
int32_t hk_prove(hk_ctx* ctx, uint64_t pk, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                 const uint64_t* kappas, uint32_t memspace, void* a_affine, void* b_affine, void* c_affine,
                 uint8_t* inf);

/* K proofs against the SAME proving key in one device pipeline (same-shape subcircuits share a key:
 * big-merkle has 5 shapes, VKD 7, VM 2 -- tree_hash_circuit.rs:192-216).  Every MSM of the batch runs
 * as one pass in which the proof index is one more bucket-set dimension, so the latency-bound
 * reduction stages are paid once per batch.  z: K x m, r, s: K, kappas: K x (num_stages-1), all
 * Montgomery Fr; outputs: K x A, K x B, K x C (affine, Montgomery), inf: K x 3 bytes. */
int32_t hk_prove_batch(hk_ctx* ctx, uint64_t pk, size_t K, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                       const uint64_t* kappas, uint32_t memspace, void* a_affine, void* b_affine, void* c_affine,
                       uint8_t* inf);

/* ---- test / measurement hooks ------------------------------------------------------- */
/* Element-wise field op on n elements (host pointers): out[i] = a[i] (op) b[i]. */
int32_t hk_field_op(hk_ctx* ctx, uint32_t field, uint32_t op, const uint64_t* a, const uint64_t* b,
                    uint64_t* out, size_t n);
/* Dependent-chain multiply throughput probe: `iters` Montgomery multiplies per thread on
 * `threads` threads; returns elapsed milliseconds (CUDA events on the context stream). */
int32_t hk_bench_field_mul(hk_ctx* ctx, uint32_t field, uint32_t threads, uint32_t iters, float* ms);
/* Per-kernel timing for the roofline report.  While enabled, the dominant kernels are bracketed
 * by CUDA events on the context stream and their work counts are recorded:
 *   HK_PROF_MSM_ACCUM_G1 / _G2  units = bucket-accumulation point additions
 *   HK_PROF_NTT_PASS            units = field elements read+written by a pass
 * hk_profile_read synchronises the stream and returns totals since the last reset.
 * on = 1: time the kernels as they run (several lanes of a proof batch overlap, so a launch's duration
 *         includes the time it shares the GPU with others);
 * on = 2: additionally serialise the lanes of hk_commit* / hk_prove* onto the context stream, so that every
 *         launch is timed alone (same kernels, same inputs; the batch itself runs slower). */
enum { HK_PROF_MSM_ACCUM_G1 = 0, HK_PROF_MSM_ACCUM_G2 = 1, HK_PROF_NTT_PASS = 2, HK_PROF_COUNT = 3 };
int32_t hk_profile_enable(hk_ctx* ctx, int32_t on);
int32_t hk_profile_reset(hk_ctx* ctx);
int32_t hk_profile_read(hk_ctx* ctx, uint32_t kernel_id, double* total_ms, uint64_t* launches, uint64_t* units);
/* Raw IMAD.WIDE.U32 issue-rate probe (the MSM roofline denominator): returns wide
 * multiply-accumulates per second. */
int32_t hk_bench_imad_wide(hk_ctx* ctx, uint32_t iters, double* macs_per_s);

#ifdef __cplusplus
}
#endif
#endif /* HEKATON_B200_H */
