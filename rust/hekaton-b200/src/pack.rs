//! Marshalling between arkworks values and the C ABI's packed limb arrays.
//!
//! `Affine<P> { x, y, infinity: bool }` is not `repr(C)`, and neither is `Fp(BigInt<N>, PhantomData)`: nothing here
//! transmutes.  Coordinates are copied limb by limb (`p.x.0 .0` is the Montgomery representation, little-endian
//! `u64` limbs -- exactly the ABI's layout), once per proving key for the bases and once per call for scalars.
use ark_ec::{
    pairing::Pairing,
    short_weierstrass::{Affine, Projective, SWCurveConfig},
};
use ark_ff::{BigInt, Fp, FpConfig, PrimeField, QuadExtConfig, QuadExtField};
use hekaton_b200_sys as sys;

/// Pairings the device library is built for.
pub trait GpuPairing: Pairing {
    const CURVE_ID: u32;
    /// u64 limbs of one base-field element (6 for BLS12-381, 4 for BN254).
    const FQ_LIMBS: usize;
    fn pack_g1(pts: &[Self::G1Affine]) -> (Vec<u64>, Vec<u8>);
    fn pack_g2(pts: &[Self::G2Affine]) -> (Vec<u64>, Vec<u8>);
    /// Jacobian (X, Y, Z) Montgomery limbs -> group element (what `hk_msm_g1` returns).
    fn unpack_g1_jacobian(limbs: &[u64]) -> Self::G1;
    fn unpack_g2_jacobian(limbs: &[u64]) -> Self::G2;
    /// Affine x || y Montgomery limbs + infinity flag -> affine point (what `hk_commit` / `hk_prove` return).
    fn unpack_g1_affine(limbs: &[u64], infinity: bool) -> Self::G1Affine;
    fn unpack_g2_affine(limbs: &[u64], infinity: bool) -> Self::G2Affine;
}

/// Montgomery limbs of scalars (`HK_SCALAR_MONT`).
pub fn pack_fr<P: FpConfig<4>>(s: &[Fp<P, 4>]) -> Vec<u64> {
    let mut out = Vec::with_capacity(4 * s.len());
    for x in s {
        out.extend_from_slice(&x.0 .0);
    }
    out
}

/// Canonical limbs of `into_bigint()` output (`HK_SCALAR_BIGINT`; `prover.rs:78-82`).
pub fn pack_bigint(s: &[BigInt<4>]) -> Vec<u64> {
    let mut out = Vec::with_capacity(4 * s.len());
    for x in s {
        out.extend_from_slice(&x.0);
    }
    out
}

pub fn unpack_fr<P: FpConfig<4>>(limbs: &[u64]) -> Vec<Fp<P, 4>> {
    limbs.chunks_exact(4).map(|c| Fp::new_unchecked(BigInt([c[0], c[1], c[2], c[3]]))).collect()
}

fn fp_from<P: FpConfig<N>, const N: usize>(limbs: &[u64]) -> Fp<P, N> {
    let mut a = [0u64; N];
    a.copy_from_slice(&limbs[..N]);
    Fp::new_unchecked(BigInt(a)) // already Montgomery: no conversion
}

/// G1-style groups: base field is a prime field with N limbs.
pub fn pack_sw_fp<C, P, const N: usize>(pts: &[Affine<C>]) -> (Vec<u64>, Vec<u8>)
where
    C: SWCurveConfig<BaseField = Fp<P, N>>,
    P: FpConfig<N>,
{
    let mut xy = Vec::with_capacity(pts.len() * 2 * N);
    let mut inf = Vec::with_capacity(pts.len());
    for p in pts {
        xy.extend_from_slice(&p.x.0 .0);
        xy.extend_from_slice(&p.y.0 .0);
        inf.push(p.infinity as u8);
    }
    (xy, inf)
}

/// G2-style groups: base field is Fq2 = c0 + c1 u over an N-limb prime field (ABI order x.c0 x.c1 y.c0 y.c1).
pub fn pack_sw_fp2<C, Q, P, const N: usize>(pts: &[Affine<C>]) -> (Vec<u64>, Vec<u8>)
where
    C: SWCurveConfig<BaseField = QuadExtField<Q>>,
    Q: QuadExtConfig<BaseField = Fp<P, N>>,
    P: FpConfig<N>,
{
    let mut xy = Vec::with_capacity(pts.len() * 4 * N);
    let mut inf = Vec::with_capacity(pts.len());
    for p in pts {
        xy.extend_from_slice(&p.x.c0.0 .0);
        xy.extend_from_slice(&p.x.c1.0 .0);
        xy.extend_from_slice(&p.y.c0.0 .0);
        xy.extend_from_slice(&p.y.c1.0 .0);
        inf.push(p.infinity as u8);
    }
    (xy, inf)
}

fn jac_fp<C, P, const N: usize>(l: &[u64]) -> Projective<C>
where
    C: SWCurveConfig<BaseField = Fp<P, N>>,
    P: FpConfig<N>,
{
    Projective::new_unchecked(fp_from::<P, N>(&l[0..]), fp_from::<P, N>(&l[N..]), fp_from::<P, N>(&l[2 * N..]))
}

fn jac_fp2<C, Q, P, const N: usize>(l: &[u64]) -> Projective<C>
where
    C: SWCurveConfig<BaseField = QuadExtField<Q>>,
    Q: QuadExtConfig<BaseField = Fp<P, N>>,
    P: FpConfig<N>,
{
    let f2 = |o: usize| QuadExtField::<Q>::new(fp_from::<P, N>(&l[o..]), fp_from::<P, N>(&l[o + N..]));
    Projective::new_unchecked(f2(0), f2(2 * N), f2(4 * N))
}

fn aff_fp<C, P, const N: usize>(l: &[u64], infinity: bool) -> Affine<C>
where
    C: SWCurveConfig<BaseField = Fp<P, N>>,
    P: FpConfig<N>,
{
    if infinity {
        Affine::identity()
    } else {
        Affine::new_unchecked(fp_from::<P, N>(&l[0..]), fp_from::<P, N>(&l[N..]))
    }
}

fn aff_fp2<C, Q, P, const N: usize>(l: &[u64], infinity: bool) -> Affine<C>
where
    C: SWCurveConfig<BaseField = QuadExtField<Q>>,
    Q: QuadExtConfig<BaseField = Fp<P, N>>,
    P: FpConfig<N>,
{
    if infinity {
        return Affine::identity();
    }
    let f2 = |o: usize| QuadExtField::<Q>::new(fp_from::<P, N>(&l[o..]), fp_from::<P, N>(&l[o + N..]));
    Affine::new_unchecked(f2(0), f2(2 * N))
}

macro_rules! impl_gpu_pairing {
    ($E:ty, $id:expr, $n:expr) => {
        impl GpuPairing for $E {
            const CURVE_ID: u32 = $id;
            const FQ_LIMBS: usize = $n;
            fn pack_g1(pts: &[Self::G1Affine]) -> (Vec<u64>, Vec<u8>) {
                pack_sw_fp(pts)
            }
            fn pack_g2(pts: &[Self::G2Affine]) -> (Vec<u64>, Vec<u8>) {
                pack_sw_fp2(pts)
            }
            fn unpack_g1_jacobian(l: &[u64]) -> Self::G1 {
                jac_fp(l)
            }
            fn unpack_g2_jacobian(l: &[u64]) -> Self::G2 {
                jac_fp2(l)
            }
            fn unpack_g1_affine(l: &[u64], infinity: bool) -> Self::G1Affine {
                aff_fp(l, infinity)
            }
            fn unpack_g2_affine(l: &[u64], infinity: bool) -> Self::G2Affine {
                aff_fp2(l, infinity)
            }
        }
    };
}
impl_gpu_pairing!(ark_bls12_381::Bls12_381, sys::HK_CURVE_BLS12_381, 6);
impl_gpu_pairing!(ark_bn254::Bn254, sys::HK_CURVE_BN254, 4);

/// Scalar-field limbs of a pairing whose `ScalarField` is a 4-limb Montgomery `Fp` (both supported curves).
/// Generic code only knows `E::ScalarField: PrimeField`; the limbs come out through `TypeId`-checked casts.
pub fn fr_mont_limbs<F: PrimeField>(s: &[F]) -> Vec<u64> {
    use core::any::TypeId;
    if TypeId::of::<F>() == TypeId::of::<ark_bls12_381::Fr>() {
        // same type: the cast only renames it
        let s = unsafe { &*(s as *const [F] as *const [ark_bls12_381::Fr]) };
        pack_fr(s)
    } else if TypeId::of::<F>() == TypeId::of::<ark_bn254::Fr>() {
        let s = unsafe { &*(s as *const [F] as *const [ark_bn254::Fr]) };
        pack_fr(s)
    } else {
        panic!("hekaton_b200 supports the scalar fields of BLS12-381 and BN254 only")
    }
}

pub fn fr_from_mont_limbs<F: PrimeField>(limbs: &[u64]) -> Vec<F> {
    use core::any::TypeId;
    if TypeId::of::<F>() == TypeId::of::<ark_bls12_381::Fr>() {
        let v: Vec<ark_bls12_381::Fr> = unpack_fr(limbs);
        let mut v = core::mem::ManuallyDrop::new(v);
        unsafe { Vec::from_raw_parts(v.as_mut_ptr() as *mut F, v.len(), v.capacity()) }
    } else if TypeId::of::<F>() == TypeId::of::<ark_bn254::Fr>() {
        let v: Vec<ark_bn254::Fr> = unpack_fr(limbs);
        let mut v = core::mem::ManuallyDrop::new(v);
        unsafe { Vec::from_raw_parts(v.as_mut_ptr() as *mut F, v.len(), v.capacity()) }
    } else {
        panic!("hekaton_b200 supports the scalar fields of BLS12-381 and BN254 only")
    }
}

pub fn curve_id_of_scalar_field<F: PrimeField>() -> u32 {
    use core::any::TypeId;
    if TypeId::of::<F>() == TypeId::of::<ark_bls12_381::Fr>() {
        sys::HK_CURVE_BLS12_381
    } else if TypeId::of::<F>() == TypeId::of::<ark_bn254::Fr>() {
        sys::HK_CURVE_BN254
    } else {
        panic!("hekaton_b200 supports the scalar fields of BLS12-381 and BN254 only")
    }
}
