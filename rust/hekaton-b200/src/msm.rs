//! `VariableBaseMSM` on the device.
//!
//! A [`BaseSet`] is a proving-key vector uploaded once (`pk.a_g`, `pk.b_g`, `pk.b_h`, `pk.h_g`,
//! `pk.ck.deltas_abc_g[i]`, `pk.deltas_g`: `cp-groth16/src/data_structures.rs:66-114`) and kept resident, with its
//! window-shifted table.  The three arkworks entry points map as follows (ark-ec 0.4 semantics):
//!
//! * `G::msm(bases, scalars)` -> `Err(min_len)` on a length mismatch  => [`BaseSet::msm_checked`]
//! * `G::msm_unchecked(bases, scalars)` -> zips to the shorter          => [`BaseSet::msm`]
//! * `G::msm_bigint(&bases[off..], bigints)`                           => [`BaseSet::msm_bigint`] with `base_offset`
//!
//! An empty MSM is the identity (`prover.rs:117` `unwrap_or(zero)`).
use crate::{
    context::Gpu,
    error::{check, Result},
    pack::{pack_bigint, GpuPairing},
};
use ark_ff::{BigInt, PrimeField};
use core::marker::PhantomData;
use hekaton_b200_sys as sys;

#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Group {
    G1,
    G2,
}

/// A base vector resident in HBM.  Freed when dropped (a `DeviceProvingKey` built from it keeps its own reference).
pub struct BaseSet<'g, E: GpuPairing> {
    gpu: &'g Gpu,
    pub(crate) handle: u64,
    pub group: Group,
    pub len: usize,
    _e: PhantomData<E>,
}

impl<'g, E: GpuPairing> BaseSet<'g, E>
where
    E::ScalarField: PrimeField<BigInt = BigInt<4>>,
{
    pub fn upload_g1(gpu: &'g Gpu, pts: &[E::G1Affine]) -> Result<Self> {
        let (xy, inf) = E::pack_g1(pts);
        let mut handle = 0u64;
        check(gpu.raw, unsafe {
            sys::hk_bases_upload_g1(gpu.raw, xy.as_ptr().cast(), inf.as_ptr(), pts.len(), sys::HK_HOST, &mut handle)
        })?;
        Ok(BaseSet { gpu, handle, group: Group::G1, len: pts.len(), _e: PhantomData })
    }

    pub fn upload_g2(gpu: &'g Gpu, pts: &[E::G2Affine]) -> Result<Self> {
        let (xy, inf) = E::pack_g2(pts);
        let mut handle = 0u64;
        check(gpu.raw, unsafe {
            sys::hk_bases_upload_g2(gpu.raw, xy.as_ptr().cast(), inf.as_ptr(), pts.len(), sys::HK_HOST, &mut handle)
        })?;
        Ok(BaseSet { gpu, handle, group: Group::G2, len: pts.len(), _e: PhantomData })
    }

    /// Build the window-shifted table (`W` x the memory; one bucket set for all windows, no final doublings).
    /// `density` = expected share of full-width scalars (witness vectors are mostly bits: ~0.15; `h`: 1.0).
    pub fn precompute(&self, density: f64) -> Result<()> {
        let permille = (density.clamp(0.001, 1.0) * 1000.0).round() as u32;
        check(self.gpu.raw, unsafe { sys::hk_bases_precompute(self.gpu.raw, self.handle, permille) })
    }

    fn run(&self, base_offset: usize, scalars: &[u64], n: usize, repr: u32, expected: Option<usize>) -> Result<Vec<u64>> {
        let width = 3 * E::FQ_LIMBS * if self.group == Group::G1 { 1 } else { 2 };
        let mut out = vec![0u64; width];
        let rc = unsafe {
            match (self.group, expected) {
                (Group::G1, Some(nb)) => sys::hk_msm_g1_checked(
                    self.gpu.raw, self.handle, base_offset, nb, scalars.as_ptr(), n, repr, sys::HK_HOST, out.as_mut_ptr().cast(),
                ),
                (Group::G1, None) => sys::hk_msm_g1(
                    self.gpu.raw, self.handle, base_offset, scalars.as_ptr(), n, repr, sys::HK_HOST, out.as_mut_ptr().cast(),
                ),
                (Group::G2, _) => sys::hk_msm_g2(
                    self.gpu.raw, self.handle, base_offset, scalars.as_ptr(), n, repr, sys::HK_HOST, out.as_mut_ptr().cast(),
                ),
            }
        };
        check(self.gpu.raw, rc)?;
        Ok(out)
    }

    /// `E::G1::msm_unchecked(&bases[base_offset..], scalars)` (G1 sets).
    pub fn msm(&self, base_offset: usize, scalars: &[E::ScalarField]) -> Result<E::G1> {
        assert_eq!(self.group, Group::G1);
        let s = crate::pack::fr_mont_limbs(scalars);
        Ok(E::unpack_g1_jacobian(&self.run(base_offset, &s, scalars.len(), sys::HK_SCALAR_MONT, None)?))
    }

    /// `E::G1::msm(bases, scalars)`: `Err` (code `HK_ERR_LENGTH`) unless `bases.len() == scalars.len()`
    /// (`committer.rs:89`; arkworks returns `Err(min_len)`).
    pub fn msm_checked(&self, scalars: &[E::ScalarField]) -> Result<E::G1> {
        assert_eq!(self.group, Group::G1);
        let s = crate::pack::fr_mont_limbs(scalars);
        Ok(E::unpack_g1_jacobian(&self.run(0, &s, scalars.len(), sys::HK_SCALAR_MONT, Some(self.len))?))
    }

    /// `G::msm_bigint(&bases[base_offset..], bigints)` as `calculate_coeff` calls it (`prover.rs:168`, offset 1).
    pub fn msm_bigint(&self, base_offset: usize, bigints: &[BigInt<4>]) -> Result<E::G1> {
        assert_eq!(self.group, Group::G1);
        let s = pack_bigint(bigints);
        Ok(E::unpack_g1_jacobian(&self.run(base_offset, &s, bigints.len(), sys::HK_SCALAR_BIGINT, None)?))
    }

    pub fn msm_bigint_g2(&self, base_offset: usize, bigints: &[BigInt<4>]) -> Result<E::G2> {
        assert_eq!(self.group, Group::G2);
        let s = pack_bigint(bigints);
        Ok(E::unpack_g2_jacobian(&self.run(base_offset, &s, bigints.len(), sys::HK_SCALAR_BIGINT, None)?))
    }
}

impl<'g, E: GpuPairing> Drop for BaseSet<'g, E> {
    fn drop(&mut self) {
        unsafe {
            sys::hk_bases_free(self.gpu.raw, self.handle);
        }
    }
}

/// `scalar_pairing(gp, scalars)` of `distributed-prover/src/pairing_ops.rs:32-39` for G1 vectors: element-wise
/// `gp[i] * scalars[i]`, normalised (the aggregator applies it to the A / C / D / prepared-input vectors of all
/// proofs: `aggregation.rs:236-242,288-310`).
pub fn scalar_pairing_g1<E: GpuPairing>(gpu: &Gpu, gp: &[E::G1Affine], scalars: &[E::ScalarField]) -> Result<Vec<E::G1Affine>> {
    assert_eq!(gp.len(), scalars.len());
    let (xy, inf) = E::pack_g1(gp);
    let s = crate::pack::fr_mont_limbs(scalars);
    let mut out = vec![0u64; xy.len()];
    let mut oinf = vec![0u8; gp.len()];
    check(gpu.raw, unsafe {
        sys::hk_scalar_mul_g1(
            gpu.raw, xy.as_ptr().cast(), inf.as_ptr(), s.as_ptr(), gp.len(), sys::HK_SCALAR_MONT, sys::HK_HOST,
            out.as_mut_ptr().cast(), oinf.as_mut_ptr(),
        )
    })?;
    let w = 2 * E::FQ_LIMBS;
    Ok((0..gp.len()).map(|i| E::unpack_g1_affine(&out[i * w..(i + 1) * w], oinf[i] != 0)).collect())
}

pub fn scalar_pairing_g2<E: GpuPairing>(gpu: &Gpu, gp: &[E::G2Affine], scalars: &[E::ScalarField]) -> Result<Vec<E::G2Affine>> {
    assert_eq!(gp.len(), scalars.len());
    let (xy, inf) = E::pack_g2(gp);
    let s = crate::pack::fr_mont_limbs(scalars);
    let mut out = vec![0u64; xy.len()];
    let mut oinf = vec![0u8; gp.len()];
    check(gpu.raw, unsafe {
        sys::hk_scalar_mul_g2(
            gpu.raw, xy.as_ptr().cast(), inf.as_ptr(), s.as_ptr(), gp.len(), sys::HK_SCALAR_MONT, sys::HK_HOST,
            out.as_mut_ptr().cast(), oinf.as_mut_ptr(),
        )
    })?;
    let w = 4 * E::FQ_LIMBS;
    Ok((0..gp.len()).map(|i| E::unpack_g2_affine(&out[i * w..(i + 1) * w], oinf[i] != 0)).collect())
}
