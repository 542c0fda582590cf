//! `hekaton-b200`: the arkworks-facing side of the B200-native backend for hekaton-system's per-subcircuit
//! CP-Groth16 ("Mirage") hot path.
//!
//! What it replaces in the reference (paths relative to the hekaton-system workspace):
//!
//! | reference call site | here |
//! |---|---|
//! | `E::G1::msm(..)` / `msm_unchecked` / `msm_bigint` at `cp-groth16/src/committer.rs:89,113`, `prover.rs:117,129,168` | [`msm::BaseSet::msm`], [`msm::BaseSet::msm_checked`], [`msm::BaseSet::msm_bigint`] |
//! | `QAP::witness_map` at `cp-groth16/src/prover.rs:123` (`LibsnarkReduction`) | [`reduction::GpuReduction`] (`impl R1CSToQAP`) |
//! | `CommitmentBuilder::{commit, prove}` (`cp-groth16/src/committer.rs:55-122`) | [`fused::GpuCommitmentBuilder`] over [`fused::DeviceProvingKey`] (one device pipeline per commitment / proof / batch of proofs) |
//! | `scalar_pairing` (`distributed-prover/src/pairing_ops.rs:32-39`) | [`msm::scalar_pairing_g1`], [`msm::scalar_pairing_g2`] |
//!
//! Layout rules (see `include/hekaton_b200.h`): field elements cross the boundary as the Montgomery limbs arkworks
//! already holds (`Fp(BigInt([u64; N]))`), copied -- never transmuted -- out of `Affine { x, y, infinity }`, which is
//! not `repr(C)`.  One [`Gpu`] context per (thread, GPU); threads proving concurrently against one key use
//! [`Gpu::fork`], which shares every uploaded resource (`mpi-snark/src/bin/node.rs:745-795`).
//!
//! There is no CPU fallback: [`Gpu::new`] fails without an sm_100 device.
//!
//! NOTE: written against arkworks 0.4 and the C ABI of this repository; the build image of this repository has no
//! Rust toolchain, so this crate is exercised only through the C ABI's Python / C++ callers there.

pub mod context;
pub mod error;
pub mod fused;
pub mod msm;
pub mod pack;
pub mod reduction;

pub use context::{with_gpu, Gpu};
pub use error::HkError;
pub use fused::{DeviceProvingKey, GpuCommitmentBuilder};
pub use pack::GpuPairing;
pub use reduction::GpuReduction;
