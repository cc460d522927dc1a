//! Drop-in for `groth16::msm::msm_curve_addition` (groth16/src/msm.rs:6) and the body of
//! `nova::PedersenCommitment::commit` (nova/src/pedersen.rs:15-20) over the C ABI in include/kgr_msm.h.
//!
//! UNTESTED: this file has never been compiled (no rustc/cargo in the development image).  The tested
//! contract is the C ABI; tests/ exercise every marshalling step below through Python/ctypes instead.
//!
//! Covered here: the two 4-limb curves of the MSM hot path.  The later rows of the same library — G2 (`KGR_CURVE_BN254_G2`,
//! 8 limbs per coordinate, needs an accessor for `Fq2`'s private array), `kgr_msm_batch`, `kgr_ntt`, `kgr_groth16_h`,
//! `kgr_r1cs_register` / `kgr_nova_cross_term` / `kgr_vec_fold` — bind the same way; INTEGRATION.md §1 and §3 list the call sites.
use core::ffi::c_char;
use std::sync::Once;

use zkstd::common::{BNAffine, BNProjective, CurveGroup};

#[allow(non_camel_case_types)]
#[repr(C)]
pub struct kgr_bases_t {
    _private: [u8; 0],
}

extern "C" {
    fn kgr_init(devices: *const i32, n_devices: i32) -> i32;
    fn kgr_last_error() -> *const c_char;
    fn kgr_bases_register(curve: i32, xy: *const u64, inf: *const u8, n: usize, out: *mut *mut kgr_bases_t) -> i32;
    fn kgr_bases_free(bases: *mut kgr_bases_t) -> i32;
    fn kgr_msm(bases: *mut kgr_bases_t, base_off: usize, scalars: *const u64, scalar_fmt: i32, n: usize, out: *mut u64) -> i32;
    fn kgr_msm_oneshot(curve: i32, xy: *const u64, inf: *const u8, n_bases: usize, scalars: *const u64, scalar_fmt: i32, n_scalars: usize,
                       out: *mut u64) -> i32;
}

const SCALARS_MONTGOMERY: i32 = 0;
static INIT: Once = Once::new();

fn check(rc: i32) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(kgr_last_error()) }.to_string_lossy().into_owned();
        panic!("kgr_msm error {rc}: {msg}"); // the reference signature is infallible (it unwraps internally)
    }
}

fn ensure_init() {
    INIT.call_once(|| check(unsafe { kgr_init(core::ptr::null(), 0) }));
}

/// Curves served by the GPU engine.  `bn_254::G2Affine` deliberately does not implement this:
/// its base field has no byte form (zkstd/src/macros/extension_field.rs:54-60) and G2 is out of scope,
/// so the prover keeps calling the reference algorithm for its two G2 MSMs (groth16/src/prover.rs:64-65).
pub trait GpuMsm: BNAffine {
    const CURVE: i32;
    fn coord_limbs(&self) -> ([u64; 4], [u64; 4]);
    fn scalar_limbs(s: &Self::Scalar) -> [u64; 4];
    fn extended_from_limbs(x: [u64; 4], y: [u64; 4], z: [u64; 4]) -> Self::Extended;
}

impl GpuMsm for bn_254::G1Affine {
    const CURVE: i32 = 0;
    fn coord_limbs(&self) -> ([u64; 4], [u64; 4]) {
        (*self.get_x().inner(), *self.get_y().inner()) // bn254/src/fq.rs:102
    }
    fn scalar_limbs(s: &bn_254::Fr) -> [u64; 4] {
        s.0 // bn254/src/fr.rs:71 (pub field, Montgomery form)
    }
    fn extended_from_limbs(x: [u64; 4], y: [u64; 4], z: [u64; 4]) -> bn_254::G1Projective {
        bn_254::G1Projective::new_unchecked(bn_254::Fq::new_unchecked(x), bn_254::Fq::new_unchecked(y), bn_254::Fq::new_unchecked(z))
    }
}

impl GpuMsm for grumpkin::Affine {
    const CURVE: i32 = 1;
    fn coord_limbs(&self) -> ([u64; 4], [u64; 4]) {
        (*self.get_x().inner(), *self.get_y().inner()) // bn254/src/fr.rs:118
    }
    fn scalar_limbs(s: &bn_254::Fq) -> [u64; 4] {
        *s.inner()
    }
    fn extended_from_limbs(x: [u64; 4], y: [u64; 4], z: [u64; 4]) -> grumpkin::Projective {
        grumpkin::Projective::new_unchecked(bn_254::Fr::new_unchecked(x), bn_254::Fr::new_unchecked(y), bn_254::Fr::new_unchecked(z))
    }
}

fn pack_points<C: GpuMsm>(bases: &[C]) -> (Vec<u64>, Vec<u8>) {
    // G1Affine is not repr(C) (bn254/src/g1.rs:17-22): copy (x, y, is_identity) into flat buffers
    let mut xy = Vec::with_capacity(8 * bases.len());
    let mut inf = Vec::with_capacity(bases.len());
    for b in bases {
        let (x, y) = b.coord_limbs();
        xy.extend_from_slice(&x);
        xy.extend_from_slice(&y);
        inf.push(b.is_identity() as u8);
    }
    (xy, inf)
}

fn pack_scalars<C: GpuMsm>(coeffs: &[C::Scalar]) -> Vec<u64> {
    let mut sc = Vec::with_capacity(4 * coeffs.len());
    for s in coeffs {
        sc.extend_from_slice(&C::scalar_limbs(s));
    }
    sc
}

fn unpack<C: GpuMsm>(out: &[u64; 12]) -> C::Extended {
    C::extended_from_limbs(out[0..4].try_into().unwrap(), out[4..8].try_into().unwrap(), out[8..12].try_into().unwrap())
}

/// Same signature and semantics as groth16/src/msm.rs:6 (pairs = zip(coeffs, bases), identity bases allowed,
/// any projective representative of the sum).
pub fn msm_curve_addition<C: GpuMsm>(bases: &[C], coeffs: &[C::Scalar]) -> C::Extended {
    ensure_init();
    let n = bases.len().min(coeffs.len());
    let (xy, inf) = pack_points(&bases[..n]);
    let sc = pack_scalars::<C>(&coeffs[..n]);
    let mut out = [0u64; 12];
    check(unsafe { kgr_msm_oneshot(C::CURVE, xy.as_ptr(), inf.as_ptr(), n, sc.as_ptr(), SCALARS_MONTGOMERY, n, out.as_mut_ptr()) });
    unpack::<C>(&out)
}

/// A base vector kept on the GPU(s): a CRS query (groth16/src/params.rs:7-29) or a Pedersen key (nova/src/pedersen.rs:6-8).
pub struct RegisteredBases<C: GpuMsm> {
    handle: *mut kgr_bases_t,
    len: usize,
    _curve: core::marker::PhantomData<C>,
}

impl<C: GpuMsm> RegisteredBases<C> {
    pub fn new(bases: &[C]) -> Self {
        ensure_init();
        let (xy, inf) = pack_points(bases);
        let mut handle = core::ptr::null_mut();
        check(unsafe { kgr_bases_register(C::CURVE, xy.as_ptr(), inf.as_ptr(), bases.len(), &mut handle) });
        Self { handle, len: bases.len(), _curve: core::marker::PhantomData }
    }

    /// sum_i coeffs[i] * bases[offset + i]   (e.g. `msm(&params.a[l..], aux)`, groth16/src/prover.rs:59)
    pub fn msm(&self, offset: usize, coeffs: &[C::Scalar]) -> C::Extended {
        let n = coeffs.len().min(self.len.saturating_sub(offset));
        let sc = pack_scalars::<C>(&coeffs[..n]);
        let mut out = [0u64; 12];
        check(unsafe { kgr_msm(self.handle, offset, sc.as_ptr(), SCALARS_MONTGOMERY, n, out.as_mut_ptr()) });
        unpack::<C>(&out)
    }

    /// PedersenCommitment::commit (nova/src/pedersen.rs:15-20)
    pub fn commit(&self, m: &[C::Scalar]) -> C {
        self.msm(0, m).to_affine()
    }
}

impl<C: GpuMsm> Drop for RegisteredBases<C> {
    fn drop(&mut self) {
        unsafe { kgr_bases_free(self.handle) };
    }
}
