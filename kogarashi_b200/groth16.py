"""G1 side of the reference's Groth16 prover on the GPU MSM engine.

Reference: groth16/src/prover.rs:49-92.  After the FFTs, `create_proof` issues six G1 MSMs
(h, l, a x2, b_g1 x2; :51-62), two G2 MSMs (:64-65, out of scope: they stay on the reference CPU path)
and then assembles
    A = r*delta + alpha + a_answer                                   (:75,81-82)
    C = r*s*delta + s*alpha + r*beta + s*a_answer + r*b1_answer + q + l   (:77,84,90-92)
Everything on the right is a linear combination of fixed CRS points, so with the CRS resident on the GPU
(`kgr_bases_register`, once per prover) A and C are ONE MSM each over the concatenated vectors
    A: [delta, alpha, a...]                       scalars [r, 1, z...]
    C: [delta, alpha, beta, a..., b_g1..., h..., l...]   scalars [r*s, s, r, s*z..., r*z..., q..., aux...]
where z = inputs ++ aux.  This is the fusion N1 asks for (pairs a_inputs/a_aux and b_g1_inputs/b_g1_aux fused,
blinding terms folded in).  Scalar products are taken mod r on the host; they are passed in canonical form.
FFT / witness generation (rows N2) and G2 (N3) are not part of this module: the caller supplies q, inputs, aux.
"""
import numpy as np

from .msm import BN254_G1, SCALARS_CANONICAL, Bases, msm_curve_addition, to_affine

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # bn254/src/fr.rs:11-16


def _scalars(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v %= FR
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


class Groth16G1Prover:
    """Holds the G1 part of `Parameters` (groth16/src/params.rs:7-29) on the GPU.

    Points are (n, 8) uint64 Montgomery arrays with (n,) uint8 infinity flags (CRS entries may be the identity,
    groth16/src/zksnark.rs:62-66,177-185)."""

    def __init__(self, delta_g1, alpha_g1, beta_g1, a, a_inf, b_g1, b_g1_inf, h, h_inf, l, l_inf):
        one = np.zeros(1, dtype=np.uint8)
        pt = lambda p: np.ascontiguousarray(p, dtype=np.uint64).reshape(1, 8)
        self.n_var = len(a)
        self.n_h, self.n_l = len(h), len(l)
        a_pts = np.concatenate([pt(delta_g1), pt(alpha_g1), a])
        a_flags = np.concatenate([one, one, a_inf])
        c_pts = np.concatenate([pt(delta_g1), pt(alpha_g1), pt(beta_g1), a, b_g1, h, l])
        c_flags = np.concatenate([one, one, one, a_inf, b_g1_inf, h_inf, l_inf])
        self.crs_a = Bases(BN254_G1, a_pts, a_flags)
        self.crs_c = Bases(BN254_G1, c_pts, c_flags)

    def commitments(self, q, inputs, aux, r, s):
        """-> (A, C) as (9,) uint64 affine [x, y, is_infinity] (Proof.a / Proof.c after `.into()`, prover.rs:94-98)."""
        z = list(inputs) + list(aux)
        assert len(z) <= self.n_var and len(q) <= self.n_h and len(aux) <= self.n_l
        zpad = z + [0] * (self.n_var - len(z))
        sa = [r, 1] + zpad
        sc = [r * s, s, r] + [s * v for v in zpad] + [r * v for v in zpad] + list(q) + [0] * (self.n_h - len(q)) + list(aux) + [0] * (self.n_l - len(aux))
        A = msm_curve_addition(self.crs_a, _scalars(sa), scalar_fmt=SCALARS_CANONICAL)
        C = msm_curve_addition(self.crs_c, _scalars(sc), scalar_fmt=SCALARS_CANONICAL)
        return to_affine(BN254_G1, A), to_affine(BN254_G1, C)

    def free(self):
        self.crs_a.free()
        self.crs_c.free()
