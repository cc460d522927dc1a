// Fixture generator for the REFERENCE side (KogarashiNetwork/Kogarashi) — not part of this repository's build.
//
// Place it inside the reference's groth16 crate (oracle/_ref/README.md has the three commands): `msm_curve_addition` is private to that
// crate (groth16/src/msm.rs:6, `mod msm;` in lib.rs), so only an in-crate test module can call it.  It prints one JSON object per case with
// the Montgomery limbs (hex, little-endian limb order) of every base, scalar and of the normalised result, which
// tests/golden/ref_vectors_to_npz.py turns into tests/golden/msm_vectors_ref.npz.
//
// Randomness: rand_xorshift::XorShiftRng with the seed the reference's own integration test uses (pallet/nova/src/tests.rs:69-74), so the
// vectors are reproducible; `G1Affine::random` / `Fr::random` are the reference's samplers (group.rs:39-41, represent.rs:80-103).
use super::msm_curve_addition;

use bn_254::{Fr, G1Affine};
use rand_xorshift::XorShiftRng;
use zkstd::common::{BNAffine, BNProjective, Group, SeedableRng, Vec};

fn hex4(l: &[u64; 4]) -> String {
    format!("[\"{:016x}\",\"{:016x}\",\"{:016x}\",\"{:016x}\"]", l[0], l[1], l[2], l[3])
}

fn emit(name: &str, points: &[G1Affine], scalars: &[Fr]) {
    let sum = msm_curve_addition(points, scalars).to_affine();
    let pts: Vec<String> = points.iter().map(|p| format!("[{},{},{}]", hex4(p.get_x().inner()), hex4(p.get_y().inner()), p.is_identity() as u8)).collect();
    let scs: Vec<String> = scalars.iter().map(|s| hex4(&s.0)).collect();
    println!(
        "{{\"name\":\"{}\",\"curve\":\"bn254_g1\",\"points\":[{}],\"scalars\":[{}],\"affine\":[{},{},{}]}}",
        name, pts.join(","), scs.join(","), hex4(sum.get_x().inner()), hex4(sum.get_y().inner()), sum.is_identity() as u8
    );
}

#[test]
fn msm_fixture() {
    let mut rng = XorShiftRng::from_seed([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5]);
    for &n in &[0usize, 1, 2, 3, 4, 31, 32, 33, 100, 1024] {
        let points = (0..n).map(|_| G1Affine::random(&mut rng)).collect::<Vec<_>>();
        let scalars = (0..n).map(|_| Fr::random(&mut rng)).collect::<Vec<_>>();
        emit(&format!("ref_uniform_{}", n), &points, &scalars);
    }
    // the special cases the CUDA path branches on: identity bases, zero / one scalars, duplicated and opposite points, fewer scalars than bases
    let n = 64;
    let mut points = (0..n).map(|_| G1Affine::random(&mut rng)).collect::<Vec<_>>();
    let mut scalars = (0..n).map(|_| Fr::random(&mut rng)).collect::<Vec<_>>();
    points[3] = G1Affine::ADDITIVE_IDENTITY;
    points[5] = points[4];
    points[7] = -points[6];
    scalars[7] = scalars[6];
    scalars[8] = Fr::zero();
    scalars[9] = Fr::one();
    scalars[10] = -Fr::one();
    emit("ref_special_64", &points, &scalars);
    emit("ref_ragged_64_50", &points, &scalars[..50]);
}
