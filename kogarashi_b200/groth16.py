"""The MSM / FFT part of the reference's Groth16 prover on the GPU engine (SURVEY.md §8f rows N1 + N2 + N3).

Reference: groth16/src/prover.rs:32-92.  `create_proof` runs seven FFTs (:36-47), six G1 MSMs (h, l, a x2, b_g1 x2;
:51-62), two G2 MSMs (b_g2 x2, :64-65) and assembles
    A = r*delta + alpha + a_answer                                          (:75,81-82)
    B = s*delta_g2 + beta_g2 + b2_answer                                    (:76,87-88)
    C = r*s*delta + s*alpha + r*beta + s*a_answer + r*b1_answer + q + l     (:77,84,90-92)
`Groth16G1Prover` computes A and C; `Groth16Prover` adds B with the G2 MSM, i.e. the whole proof.
Here the G1 CRS (`Parameters::{h, l, a, b_g1}`, groth16/src/params.rs:7-29) is registered on the GPU once per prover;
per proof the H coefficients come from the device NTT (kogarashi_b200.fft) and the `a_inputs / a_aux` and
`b_g1_inputs / b_g1_aux` pairs are fused into one MSM each over z = inputs ++ aux (the prover splits them only because
of its slice layout, :58-62).  The blinding terms are a five-point MSM, so no curve arithmetic is done in Python.
"""
import numpy as np

from .fft import Fft
from .msm import BN254_G1, BN254_G2, SCALARS_CANONICAL, SCALARS_MONTGOMERY, Bases, groth16_msms, msm_batch, msm_curve_addition, proj_add, to_affine

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # bn254/src/fr.rs:11-16


class ProverSubVersionCrsAttack(ValueError):
    """groth16/src/error.rs:3 — `create_proof` refuses a CRS whose delta_g1 or delta_g2 is the identity (prover.rs:67-69)."""


def _vk_points(points, inf, words):
    """vk elements as a (k, words) array + (k,) infinity flags.  The reference encodes the identity as (0, R, is_infinity = true)
    (zkstd/src/macros/curve/weierstrass/group.rs:106-110); (0, 0) — the device's own identity encoding — is accepted as well."""
    pts = np.stack([np.ascontiguousarray(p, dtype=np.uint64).reshape(words) for p in points])
    flags = np.zeros(len(points), dtype=np.uint8) if inf is None else np.ascontiguousarray(inf, dtype=np.uint8).reshape(len(points)).copy()
    flags |= (~pts.any(axis=1)).astype(np.uint8)
    return pts, flags


def _canonical(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v %= FR
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


class Groth16G1Prover:
    """Points are (n, 8) uint64 Montgomery arrays with (n,) uint8 infinity flags (CRS entries may be the identity,
    groth16/src/zksnark.rs:62-66,177-185)."""

    def __init__(self, delta_g1, alpha_g1, beta_g1, a, a_inf, b_g1, b_g1_inf, h, h_inf, l, l_inf, precompute=False, vk_inf=None):
        """vk_inf: optional is_infinity flags of (delta_g1, alpha_g1, beta_g1).  An identity delta_g1 raises ProverSubVersionCrsAttack
        before anything is uploaded — the reference returns that error from create_proof (prover.rs:67-69)."""
        self.vk, self.vk_inf = _vk_points([delta_g1, alpha_g1, beta_g1], vk_inf, 8)
        if self.vk_inf[0]:
            raise ProverSubVersionCrsAttack("vk.delta_g1 is the identity")
        self.a, self.b_g1, self.h, self.l = (Bases(BN254_G1, p, f) for p, f in ((a, a_inf), (b_g1, b_g1_inf), (h, h_inf), (l, l_inf)))
        self.vk_bases = Bases(BN254_G1, self.vk, self.vk_inf)
        if precompute:
            for b in (self.a, self.b_g1, self.h, self.l):
                b.precompute(0)

    def _g1_jobs(self, q, inputs, aux, r, scalar_fmt):
        """The G1 queries of prover.rs:51-62 as one batch (independent MSMs overlap on the device) + the blinding of A."""
        z = np.concatenate([np.asarray(inputs, dtype=np.uint64).reshape(-1, 4), np.asarray(aux, dtype=np.uint64).reshape(-1, 4)])
        return z, [(self.a, z, 0, scalar_fmt),                      # a_inputs + a_aux       (:58-59, :80)
                   (self.b_g1, z, 0, scalar_fmt),                   # b_g1_inputs + b_g1_aux (:61-62, :86)
                   (self.h, q, 0, scalar_fmt),                      # :51
                   (self.l, aux, 0, scalar_fmt),                    # :56
                   (self.vk_bases, _canonical([r, 1]), 0, SCALARS_CANONICAL)]   # r*delta + alpha (:75)

    def _assemble_g1(self, res, r, s):
        a_answer, b1_answer, q_pt, l_pt, blind_a = res[:5]
        delta, alpha, beta = self.vk
        g_a = proj_add(BN254_G1, blind_a, a_answer)
        aa, ba = to_affine(BN254_G1, a_answer), to_affine(BN254_G1, b1_answer)
        pts = np.stack([aa[:8], ba[:8], delta, alpha, beta])
        inf = np.array([aa[8], ba[8], self.vk_inf[0], self.vk_inf[1], self.vk_inf[2]], dtype=np.uint8)
        blind_c = msm_curve_addition(pts, _canonical([s, r, r * s, s, r]), curve=BN254_G1, inf=inf, scalar_fmt=SCALARS_CANONICAL)
        g_c = proj_add(BN254_G1, proj_add(BN254_G1, blind_c, q_pt), l_pt)
        return to_affine(BN254_G1, g_a), to_affine(BN254_G1, g_c)

    def prove_g1(self, q, inputs, aux, r, s, scalar_fmt=SCALARS_MONTGOMERY):
        """q, inputs, aux: (len, 4) uint64 scalars in `scalar_fmt`; r, s: Python ints (the prover's blinding factors, prover.rs:71-72).
        -> (A, C) as (9,) uint64 affine [x, y, is_infinity] = Proof.a / Proof.c after `.into()` (prover.rs:94-98)."""
        _, jobs = self._g1_jobs(q, inputs, aux, r, scalar_fmt)
        return self._assemble_g1(msm_batch(jobs), r, s)

    def prove_from_evaluations(self, log_n, a_evals, b_evals, c_evals, inputs, aux, r, s):
        """The whole G1 side of create_proof from the R1CS evaluations (prover.rs:33): one kgr_groth16_msms call — H on the device feeding the h
        query directly, the other queries overlapped with it.  All scalars are Montgomery arrays (the reference's in-memory form)."""
        _, jobs = self._g1_jobs(np.zeros((0, 4), dtype=np.uint64), inputs, aux, r, SCALARS_MONTGOMERY)
        del jobs[2]                                                   # the h query is the fused part of the call
        q_pt, q, res = groth16_msms(log_n, a_evals, b_evals, c_evals, self.h, jobs)
        return self._assemble_g1([res[0], res[1], q_pt, res[2], res[3]], r, s) + (q,)

    def commitments(self, q, inputs, aux, r, s):
        """Same as prove_g1 for scalars given as Python integers."""
        return self.prove_g1(_canonical(list(q)), _canonical(list(inputs)), _canonical(list(aux)), r, s, scalar_fmt=SCALARS_CANONICAL)

    def free(self):
        for b in (self.a, self.b_g1, self.h, self.l, self.vk_bases):
            b.free()


class Groth16Prover(Groth16G1Prover):
    """All three proof elements.  G2 points are (n, 16) uint64 = x.c0 x.c1 y.c0 y.c1 Montgomery (bn254/src/g2.rs:15-21)."""

    def __init__(self, delta_g1, alpha_g1, beta_g1, a, a_inf, b_g1, b_g1_inf, h, h_inf, l, l_inf, delta_g2, beta_g2, b_g2, b_g2_inf, precompute=False,
                 vk_inf=None, vk_g2_inf=None):
        """vk_g2_inf: optional is_infinity flags of (delta_g2, beta_g2); an identity delta_g2 raises ProverSubVersionCrsAttack (prover.rs:67-69)."""
        self.vk_g2, self.vk_g2_inf = _vk_points([delta_g2, beta_g2], vk_g2_inf, 16)
        if self.vk_g2_inf[0]:
            raise ProverSubVersionCrsAttack("vk.delta_g2 is the identity")
        super().__init__(delta_g1, alpha_g1, beta_g1, a, a_inf, b_g1, b_g1_inf, h, h_inf, l, l_inf, precompute=precompute, vk_inf=vk_inf)
        self.b_g2 = Bases(BN254_G2, b_g2, b_g2_inf)
        self.vk_g2_bases = Bases(BN254_G2, self.vk_g2, self.vk_g2_inf)
        if precompute:
            self.b_g2.precompute(0)

    def prove(self, q, inputs, aux, r, s, scalar_fmt=SCALARS_MONTGOMERY):
        """-> (A, B, C): A, C (9,) uint64 and B (17,) uint64 affine [x, y, is_infinity] = Proof {a, b, c} (prover.rs:94-98).
        All eight MSMs of prover.rs:51-65 (pairs fused) and the two independent blinding sums go to the device as one batch."""
        z, jobs = self._g1_jobs(q, inputs, aux, r, scalar_fmt)
        # the G2 query is the longest job (3x the arithmetic per point): it goes first so that the G1 queries overlap with it
        jobs = [(self.b_g2, z, 0, scalar_fmt)] + jobs + [(self.vk_g2_bases, _canonical([s, 1]), 0, SCALARS_CANONICAL)]   # :64-65,:87 / :76
        res = msm_batch(jobs)
        g_a, g_c = self._assemble_g1(res[1:6], r, s)
        return g_a, to_affine(BN254_G2, proj_add(BN254_G2, res[6], res[0])), g_c

    def prove_from_evaluations(self, log_n, a_evals, b_evals, c_evals, inputs, aux, r, s):
        """create_proof after witness generation (prover.rs:33-98) as one kgr_groth16_msms call + the dependent blinding sum."""
        z, jobs = self._g1_jobs(np.zeros((0, 4), dtype=np.uint64), inputs, aux, r, SCALARS_MONTGOMERY)
        del jobs[2]
        jobs = [(self.b_g2, z, 0, SCALARS_MONTGOMERY)] + jobs + [(self.vk_g2_bases, _canonical([s, 1]), 0, SCALARS_CANONICAL)]
        q_pt, q, res = groth16_msms(log_n, a_evals, b_evals, c_evals, self.h, jobs)
        g_a, g_c = self._assemble_g1([res[1], res[2], q_pt, res[3], res[4]], r, s)
        return g_a, to_affine(BN254_G2, proj_add(BN254_G2, res[5], res[0])), g_c, q

    def proof(self, q, inputs, aux, r, s):
        """Scalars as Python integers."""
        return self.prove(_canonical(list(q)), _canonical(list(inputs)), _canonical(list(aux)), r, s, scalar_fmt=SCALARS_CANONICAL)

    def free(self):
        super().free()
        self.b_g2.free()
        self.vk_g2_bases.free()
