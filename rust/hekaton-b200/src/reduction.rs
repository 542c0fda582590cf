//! `GpuReduction`: an `R1CSToQAP` whose witness map runs on the device.
//!
//! Plug point: `CPGroth16<E, QAP>` / `CommitmentBuilder<_, E, QAP>` take the reduction as a type
//! parameter (`cp-groth16/src/lib.rs:18`, `committer.rs:15-37`; chosen at `distributed-prover/src/util.rs:20`).
//! After the one-token fix of `committer.rs:107` (patches/0001) every FFT of the prover
//! (3 iFFT + 3 coset FFT + pointwise (ab - c) / Z_H + 1 coset iFFT, ark-groth16 `witness_map_from_matrices`) and
//! the sparse products A z, B z, C z run in `hk_qap_witness_map`.
//!
//! `instance_map_with_evaluation` and `h_query_scalars` are setup-time functions and simply delegate to
//! `LibsnarkReduction`, so keys generated with either reduction are identical.
use crate::{context::with_gpu, error::check, pack};
use ark_ff::PrimeField;
use ark_groth16::r1cs_to_qap::{LibsnarkReduction, R1CSToQAP};
use ark_poly::EvaluationDomain;
use ark_relations::r1cs::{ConstraintMatrices, ConstraintSystemRef, Matrix, SynthesisError};
use hekaton_b200_sys as sys;
use std::{cell::RefCell, collections::HashMap};

pub struct GpuReduction;

/// CSR form of one constraint matrix: row pointers, column indices (over [instance || witness]), Montgomery values.
pub(crate) struct Csr {
    pub row_ptr: Vec<u32>,
    pub col: Vec<u32>,
    pub val: Vec<u64>,
}

pub(crate) fn to_csr<F: PrimeField>(m: &Matrix<F>) -> Csr {
    let nnz: usize = m.iter().map(|r| r.len()).sum();
    let mut row_ptr = Vec::with_capacity(m.len() + 1);
    let mut col = Vec::with_capacity(nnz);
    let mut coeffs = Vec::with_capacity(nnz);
    row_ptr.push(0u32);
    for row in m {
        for (c, j) in row {
            col.push(*j as u32);
            coeffs.push(*c);
        }
        row_ptr.push(col.len() as u32);
    }
    Csr { row_ptr, col, val: pack::fr_mont_limbs(&coeffs) }
}

/// Upload the three matrices of a constraint system; returns the `hk_r1cs_upload` handle.
pub(crate) fn upload_matrices<F: PrimeField>(
    gpu: &crate::Gpu,
    matrices: &ConstraintMatrices<F>,
    num_inputs: usize,
    num_constraints: usize,
) -> crate::error::Result<u64> {
    let (a, b, c) = (to_csr(&matrices.a), to_csr(&matrices.b), to_csr(&matrices.c));
    let mut h = 0u64;
    check(gpu.raw, unsafe {
        sys::hk_r1cs_upload(
            gpu.raw,
            num_constraints,
            num_inputs,
            matrices.num_witness_variables,
            a.row_ptr.as_ptr(), a.col.as_ptr(), a.val.as_ptr(),
            b.row_ptr.as_ptr(), b.col.as_ptr(), b.val.as_ptr(),
            c.row_ptr.as_ptr(), c.col.as_ptr(), c.val.as_ptr(),
            &mut h,
        )
    })?;
    Ok(h)
}

/// Content fingerprint of a constraint system's matrices (64-bit FNV-1a over shape, columns and coefficients).
/// A worker sees a handful of shapes (big-merkle 5, VKD 7, VM 2) many times; the fingerprint finds the resident
/// copy.  Cost: one pass over the non-zeros, far below the upload it avoids.
fn fingerprint<F: PrimeField>(m: &ConstraintMatrices<F>, num_inputs: usize, num_constraints: usize) -> u64 {
    let mut h = 0xcbf29ce484222325u64;
    let mut eat = |x: u64| {
        h ^= x;
        h = h.wrapping_mul(0x100000001b3);
    };
    eat(num_inputs as u64);
    eat(num_constraints as u64);
    eat(m.num_witness_variables as u64);
    for mat in [&m.a, &m.b, &m.c] {
        eat(mat.len() as u64);
        for row in mat.iter() {
            eat(row.len() as u64);
            for (c, j) in row {
                eat(*j as u64);
                for limb in c.into_bigint().as_ref() {
                    eat(*limb);
                }
            }
        }
    }
    h
}

thread_local! {
    /// (curve id, fingerprint) -> resident r1cs handle.  Handles are shared by all forks of the root context.
    static SHAPES: RefCell<HashMap<(u32, u64), u64>> = RefCell::new(HashMap::new());
}

impl GpuReduction {
    /// Make the matrices of `cs` resident ahead of time (e.g. right after `ProvingKeys::deserialize`).
    pub fn preload<F: PrimeField>(cs: ConstraintSystemRef<F>) -> Result<u64, SynthesisError> {
        let matrices = cs.to_matrices().ok_or(SynthesisError::AssignmentMissing)?;
        Ok(Self::resident(&matrices, cs.num_instance_variables(), cs.num_constraints()))
    }

    fn resident<F: PrimeField>(matrices: &ConstraintMatrices<F>, num_inputs: usize, num_constraints: usize) -> u64 {
        let curve = pack::curve_id_of_scalar_field::<F>();
        let key = (curve, fingerprint(matrices, num_inputs, num_constraints));
        SHAPES.with(|s| {
            if let Some(h) = s.borrow().get(&key) {
                return *h;
            }
            let h = with_gpu(curve, |g| upload_matrices(g, matrices, num_inputs, num_constraints))
                .expect("hk_r1cs_upload failed");
            s.borrow_mut().insert(key, h);
            h
        })
    }
}

impl R1CSToQAP for GpuReduction {
    #[inline]
    fn instance_map_with_evaluation<F: PrimeField, D: EvaluationDomain<F>>(
        cs: ConstraintSystemRef<F>,
        t: &F,
    ) -> Result<(Vec<F>, Vec<F>, Vec<F>, F, usize, usize), SynthesisError> {
        LibsnarkReduction::instance_map_with_evaluation::<F, D>(cs, t)
    }

    /// ark-groth16 `LibsnarkReduction::witness_map_from_matrices`, on the device: returns all
    /// `N = next_pow2(num_constraints + num_inputs)` coefficients of `h` (`prover.rs:123-129` uses `h[..N-1]`).
    fn witness_map_from_matrices<F: PrimeField, D: EvaluationDomain<F>>(
        matrices: &ConstraintMatrices<F>,
        num_inputs: usize,
        num_constraints: usize,
        full_assignment: &[F],
    ) -> Result<Vec<F>, SynthesisError> {
        let domain = D::new(num_constraints + num_inputs).ok_or(SynthesisError::PolynomialDegreeTooLarge)?;
        let n = domain.size();
        let curve = pack::curve_id_of_scalar_field::<F>();
        let r1cs = Self::resident(matrices, num_inputs, num_constraints);
        let z = pack::fr_mont_limbs(full_assignment);
        let mut h = vec![0u64; 4 * n];
        with_gpu(curve, |g| {
            check(g.raw, unsafe { sys::hk_qap_witness_map(g.raw, r1cs, z.as_ptr(), sys::HK_HOST, h.as_mut_ptr()) })
        })
        .map_err(|_| SynthesisError::PolynomialDegreeTooLarge)?;
        Ok(pack::fr_from_mont_limbs(&h))
    }

    #[inline]
    fn h_query_scalars<F: PrimeField, D: EvaluationDomain<F>>(
        max_power: usize,
        t: F,
        zt: F,
        delta_inverse: F,
    ) -> Result<Vec<F>, SynthesisError> {
        LibsnarkReduction::h_query_scalars::<F, D>(max_power, t, zt, delta_inverse)
    }
}
