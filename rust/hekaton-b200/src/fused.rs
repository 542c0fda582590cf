//! The fused per-subcircuit path: one device pipeline per commitment / proof / batch of proofs.
//!
//! [`DeviceProvingKey`] makes a `ark_cp_groth16::ProvingKey` resident once (`hk_pk_create`: base vectors, window
//! tables, constraint matrices).  [`GpuCommitmentBuilder`] has the reference's `CommitmentBuilder` interface
//! (`cp-groth16/src/committer.rs:15-122`: `new`, `commit`, `prove`) and the same division of labour: the host
//! synthesises constraints and draws randomness in the reference's order (kappa in `commit`, `committer.rs:85`;
//! r then s in `prove`, `prover.rs:28-29`), the device does everything from the assignment on:
//!
//! * `commit`: `MSM(ck[stage], w_stage) + kappa * last_delta_g` -> affine            (`committer.rs:85-91`)
//! * `prove`:  A, B (G1 and G2), L, witness map, H, `C = sA + rB - rs delta + L + H - sum kappa_i delta_i`,
//!             three `into_affine`                                    (`prover.rs:77-154`, `committer.rs:106-114`)
//!
//! The device never sees an RNG, so proofs are byte-identical to the reference's under the same seeded host RNG.
use crate::{
    context::Gpu,
    error::{check, Result},
    pack::{self, GpuPairing},
    reduction::upload_matrices,
};
use ark_cp_groth16::{
    data_structures::{Comm, CommRandomness, Proof, ProvingKey},
    MultiStageConstraintSynthesizer, MultiStageConstraintSystem,
};
use ark_ff::{BigInt, PrimeField, UniformRand};
use ark_relations::r1cs::{ConstraintMatrices, OptimizationGoal, SynthesisError};
use ark_std::rand::Rng;
use core::{marker::PhantomData, ptr};
use hekaton_b200_sys as sys;

/// A proving key resident on one GPU (shared by all forks of the context that created it).
pub struct DeviceProvingKey<'g, E: GpuPairing> {
    gpu: &'g Gpu,
    handle: u64,
    pub num_stages: usize,
    pub num_variables: usize,
    _e: PhantomData<E>,
}

fn up_g1<E: GpuPairing>(gpu: &Gpu, pts: &[E::G1Affine]) -> Result<u64> {
    let (xy, inf) = E::pack_g1(pts);
    let mut h = 0u64;
    check(gpu.raw, unsafe { sys::hk_bases_upload_g1(gpu.raw, xy.as_ptr().cast(), inf.as_ptr(), pts.len(), sys::HK_HOST, &mut h) })?;
    Ok(h)
}

fn up_g2<E: GpuPairing>(gpu: &Gpu, pts: &[E::G2Affine]) -> Result<u64> {
    let (xy, inf) = E::pack_g2(pts);
    let mut h = 0u64;
    check(gpu.raw, unsafe { sys::hk_bases_upload_g2(gpu.raw, xy.as_ptr().cast(), inf.as_ptr(), pts.len(), sys::HK_HOST, &mut h) })?;
    Ok(h)
}

impl<'g, E: GpuPairing> DeviceProvingKey<'g, E>
where
    E::ScalarField: PrimeField<BigInt = BigInt<4>>,
{
    /// `matrices` / `num_inputs` / `num_constraints`: the finalised constraint system of this subcircuit shape
    /// (`cs.to_matrices()` after the last stage); `stage_ranges`: `MultiStageConstraintSystem::variable_range_for_stage`.
    pub fn new(
        gpu: &'g Gpu,
        pk: &ProvingKey<E>,
        matrices: &ConstraintMatrices<E::ScalarField>,
        num_inputs: usize,
        num_constraints: usize,
        stage_ranges: &[core::ops::Range<usize>],
    ) -> Result<Self> {
        assert!(!stage_ranges.is_empty() && stage_ranges.len() <= 8, "1..8 stages");
        assert_eq!(stage_ranges.len(), pk.ck.deltas_abc_g.len());
        let r1cs = upload_matrices(gpu, matrices, num_inputs, num_constraints)?;
        let handles_g1 = [up_g1::<E>(gpu, &pk.a_g)?, up_g1::<E>(gpu, &pk.b_g)?, up_g1::<E>(gpu, &pk.h_g)?];
        let b_h = up_g2::<E>(gpu, &pk.b_h)?;
        let deltas = up_g1::<E>(gpu, &pk.deltas_g)?;
        let mut ck = [0u64; 8];
        for (i, v) in pk.ck.deltas_abc_g.iter().enumerate() {
            ck[i] = up_g1::<E>(gpu, v)?;
        }
        let (alpha_g, _) = E::pack_g1(&[pk.vk.alpha_g]);
        let (beta_g, _) = E::pack_g1(&[pk.beta_g]);
        let (beta_h, _) = E::pack_g2(&[pk.vk.beta_h]);
        let (last_delta_h, _) = E::pack_g2(&[pk.last_delta_h()]);
        let mut stage_start = [0u32; 9];
        for (i, r) in stage_ranges.iter().enumerate() {
            stage_start[i] = r.start as u32;
            stage_start[i + 1] = r.end as u32;
        }
        let desc = sys::hk_pk_desc {
            a_g: handles_g1[0],
            b_g: handles_g1[1],
            h_g: handles_g1[2],
            b_h,
            ck,
            num_stages: stage_ranges.len() as u32,
            deltas_g: deltas,
            alpha_g: alpha_g.as_ptr().cast(),
            beta_g: beta_g.as_ptr().cast(),
            beta_h: beta_h.as_ptr().cast(),
            last_delta_h: last_delta_h.as_ptr().cast(),
            r1cs,
            stage_start,
            witness_density_permille: 0,
        };
        let mut handle = 0u64;
        check(gpu.raw, unsafe { sys::hk_pk_create(gpu.raw, &desc, &mut handle) })?;
        // the key keeps what it needs (extended copies + the shared h_g set); the plain uploads can go
        for h in [handles_g1[0], handles_g1[1], b_h, deltas].into_iter().chain(ck.into_iter().take(stage_ranges.len())) {
            unsafe { sys::hk_bases_free(gpu.raw, h) };
        }
        Ok(DeviceProvingKey {
            gpu,
            handle,
            num_stages: stage_ranges.len(),
            num_variables: num_inputs + matrices.num_witness_variables,
            _e: PhantomData,
        })
    }

    /// `committer.rs:85-91`.  `z`: the assignment so far (instance, then every witness up to this stage).
    pub fn commit(&self, gpu: &Gpu, stage: usize, z: &[E::ScalarField], kappa: E::ScalarField) -> Result<E::G1Affine> {
        let zl = pack::fr_mont_limbs(z);
        let kl = pack::fr_mont_limbs(&[kappa]);
        let mut out = vec![0u64; 2 * E::FQ_LIMBS];
        let mut inf = 0u8;
        check(gpu.raw, unsafe {
            sys::hk_commit(gpu.raw, self.handle, stage as u32, zl.as_ptr(), kl.as_ptr(), sys::HK_HOST, out.as_mut_ptr().cast(), &mut inf)
        })?;
        Ok(E::unpack_g1_affine(&out, inf != 0))
    }

    /// `prover.rs:53-156` + `committer.rs:106-114` for one subcircuit: returns (A, B, C).
    pub fn prove(
        &self,
        gpu: &Gpu,
        z: &[E::ScalarField],
        r: E::ScalarField,
        s: E::ScalarField,
        kappas: &[E::ScalarField],
    ) -> Result<(E::G1Affine, E::G2Affine, E::G1Affine)> {
        let mut v = self.prove_batch(gpu, &[z], &[r], &[s], &[kappas])?;
        Ok(v.pop().unwrap())
    }

    /// K same-shape subcircuits in ONE device pipeline (`hk_prove_batch`): a worker that holds several subcircuits
    /// of one shape (`mpi-snark/src/bin/node.rs:548-551`) proves them together.
    pub fn prove_batch(
        &self,
        gpu: &Gpu,
        zs: &[&[E::ScalarField]],
        rs: &[E::ScalarField],
        ss: &[E::ScalarField],
        kappas: &[&[E::ScalarField]],
    ) -> Result<Vec<(E::G1Affine, E::G2Affine, E::G1Affine)>> {
        let k = zs.len();
        assert!(rs.len() == k && ss.len() == k && kappas.len() == k);
        let mut zl = Vec::with_capacity(4 * k * self.num_variables);
        for z in zs {
            assert_eq!(z.len(), self.num_variables, "full assignment (instance || witness) expected");
            zl.extend_from_slice(&pack::fr_mont_limbs(z));
        }
        let rl = pack::fr_mont_limbs(rs);
        let sl = pack::fr_mont_limbs(ss);
        let mut kl = Vec::with_capacity(4 * k * (self.num_stages - 1));
        for kk in kappas {
            assert_eq!(kk.len(), self.num_stages - 1, "one kappa per committed stage (committer.rs:111)");
            kl.extend_from_slice(&pack::fr_mont_limbs(kk));
        }
        let (w1, w2) = (2 * E::FQ_LIMBS, 4 * E::FQ_LIMBS);
        let mut a = vec![0u64; k * w1];
        let mut b = vec![0u64; k * w2];
        let mut c = vec![0u64; k * w1];
        let mut inf = vec![0u8; 3 * k];
        check(gpu.raw, unsafe {
            sys::hk_prove_batch(
                gpu.raw,
                self.handle,
                k,
                zl.as_ptr(),
                rl.as_ptr(),
                sl.as_ptr(),
                if kl.is_empty() { ptr::null() } else { kl.as_ptr() },
                sys::HK_HOST,
                a.as_mut_ptr().cast(),
                b.as_mut_ptr().cast(),
                c.as_mut_ptr().cast(),
                inf.as_mut_ptr(),
            )
        })?;
        Ok((0..k)
            .map(|i| {
                (
                    E::unpack_g1_affine(&a[i * w1..(i + 1) * w1], inf[3 * i] != 0),
                    E::unpack_g2_affine(&b[i * w2..(i + 1) * w2], inf[3 * i + 1] != 0),
                    E::unpack_g1_affine(&c[i * w1..(i + 1) * w1], inf[3 * i + 2] != 0),
                )
            })
            .collect())
    }
}

impl<'g, E: GpuPairing> Drop for DeviceProvingKey<'g, E> {
    fn drop(&mut self) {
        unsafe {
            sys::hk_pk_free(self.gpu.raw, self.handle);
        }
    }
}

/// Drop-in for `ark_cp_groth16::committer::CommitmentBuilder` (same constructor and method names, same error
/// behaviour, same RNG draw order) that runs the arithmetic on the device.
pub struct GpuCommitmentBuilder<'a, 'g, C, E>
where
    C: MultiStageConstraintSynthesizer<E::ScalarField>,
    E: GpuPairing,
{
    pub cs: MultiStageConstraintSystem<E::ScalarField>,
    pub circuit: C,
    cur_stage: usize,
    gpu: &'a Gpu,
    pk: &'a DeviceProvingKey<'g, E>,
}

impl<'a, 'g, C, E> GpuCommitmentBuilder<'a, 'g, C, E>
where
    C: MultiStageConstraintSynthesizer<E::ScalarField>,
    E: GpuPairing,
    E::ScalarField: PrimeField<BigInt = BigInt<4>>,
{
    pub fn new(circuit: C, gpu: &'a Gpu, pk: &'a DeviceProvingKey<'g, E>) -> Self {
        let mut mscs = MultiStageConstraintSystem::default();
        mscs.set_optimization_goal(OptimizationGoal::Constraints);
        Self { cs: mscs, circuit, cur_stage: 0, gpu, pk }
    }

    /// `CommitmentBuilder::commit` (`committer.rs:55-98`).
    pub fn commit(&mut self, rng: &mut impl Rng) -> core::result::Result<(Comm<E>, CommRandomness<E>), SynthesisError> {
        self.circuit.generate_constraints(self.cur_stage, &mut self.cs)?;
        assert!(self.cur_stage < self.pk.num_stages, "no more values left in committing key");
        let randomness = E::ScalarField::rand(rng);
        let z = self.cs.full_assignment();
        let commitment =
            self.pk.commit(self.gpu, self.cur_stage, &z, randomness).map_err(|_| SynthesisError::AssignmentMissing)?;
        self.cur_stage += 1;
        Ok((commitment, randomness))
    }

    /// `CommitmentBuilder::prove` (`committer.rs:100-122`): last-stage synthesis, `cs.finalize()`, r and s from
    /// `rng` (in that order, `prover.rs:28-29`), then the fused device proof.
    pub fn prove(
        mut self,
        comms: &[Comm<E>],
        comm_rands: &[CommRandomness<E>],
        rng: &mut impl Rng,
    ) -> core::result::Result<Proof<E>, SynthesisError> {
        let r = E::ScalarField::rand(rng);
        let s = E::ScalarField::rand(rng);
        self.circuit.generate_constraints(self.circuit.last_stage(), &mut self.cs)?;
        self.cs.finalize();
        assert_eq!(self.pk.num_stages, comm_rands.len() + 1);
        let z = self.cs.full_assignment();
        let (a, b, c) = self.pk.prove(self.gpu, &z, r, s, comm_rands).map_err(|_| SynthesisError::AssignmentMissing)?;
        Ok(Proof { a, b, c, ds: comms.to_vec() })
    }
}
