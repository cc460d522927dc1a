"""BASELINE config #4: Groth16 proof of the reference's example circuit (groth16/examples/simple.rs, x^3 + x + 5 = 35)
under fixed randomness.  CPU part: the big-int restatement of setup + create_proof yields a proof that satisfies the
Groth16 equation (checked in the exponent).  GPU part: A, C (G1) and B (G2) computed by the CUDA MSM engine from the
device-resident CRS are byte-identical to the all-CPU computation, i.e. the whole 259-byte proof."""
import numpy as np
import pytest

from oracle import groth16_ref as G
from oracle import pyref as B

SETUP_SEED = bytes.fromhex("5962be5d763d318d17db37325406bce5")   # pallet/nova/src/tests.rs:69-74
PROVE_SEED = bytes(range(1, 17))


@pytest.fixture(scope="module")
def fixture():
    P, trap, uvw = G.setup(B.XorShift128(SETUP_SEED))
    proof, internals = G.create_proof(P, B.XorShift128(PROVE_SEED), 3, 35)
    return P, trap, uvw, proof, internals


def test_example_circuit_shape():
    cs = G.example_circuit(3, 35)
    assert (cs.m, cs.l, cs.m_l_1) == (4, 3, 3)
    assert cs.x == [1, 3, 35] and cs.w == [9, 27, 30]
    a, b, c = cs.evaluate()
    assert all((ai * bi - ci) % G.R == 0 for ai, bi, ci in zip(a, b, c))


def test_fft_roundtrip_and_coset():
    f = G.Fft(2)
    c = [5, 7, 11, 13]
    assert f.idft(f.dft(c)) == c and f.coset_idft(f.coset_dft(c)) == c
    # dft evaluates the polynomial on the domain
    assert f.dft(c)[1] == sum(cj * pow(f.omega, j, G.R) for j, cj in enumerate(c)) % G.R


def test_reference_prover_restatement_verifies(fixture):
    P, trap, uvw, proof, internals = fixture
    assert len(P["h"]) == 3 and len(P["a"]) == 6 and len(P["l"]) == 3 and len(P["ic"]) == 3
    assert len(internals["q"]) <= 3
    ok_points, ok_pairing = G.exponent_check(trap, uvw, internals, proof)
    assert ok_points and ok_pairing
    assert len(G.proof_bytes(proof)) == 65 + 129 + 65
    # a wrong public output must not verify
    bad, bad_int = G.create_proof(P, B.XorShift128(PROVE_SEED), 3, 36)
    assert not G.exponent_check(trap, uvw, bad_int, bad)[1]


def test_closed_form_lagrange_and_chain_circuit():
    f = G.Fft(3)
    tau = 123456789123456789
    by_idft = f.idft([pow(tau, i, G.R) for i in range(8)])
    assert G.lagrange_at(8, f.omega, tau) == by_idft + [0] * (8 - len(by_idft))
    cs, out = G.chain_circuit(5, 3)
    assert (cs.m, cs.l, cs.m_l_1) == (16, 3, 15)
    a, b, c = cs.evaluate()
    assert all((ai * bi - ci) % G.R == 0 for ai, bi, ci in zip(a, b, c))
    # one step of the chain is the reference's example circuit
    cs1, out1 = G.chain_circuit(1, 3)
    ex = G.example_circuit(3, 35)
    assert out1 == 35 and (cs1.a, cs1.b, cs1.c, cs1.x, cs1.w) == (ex.a, ex.b, ex.c, ex.x, ex.w)
    # exponent bookkeeping agrees with the point-based setup on the example
    E, trap, uvw = G.crs_exponents(ex, B.XorShift128(SETUP_SEED))
    P, trap2, uvw2 = G.setup(B.XorShift128(SETUP_SEED))
    assert trap == trap2 and uvw == uvw2
    assert [G.G1.mul(G.G1.g, e) if e else None for e in E["a"]] == P["a"] and [G.G1.mul(G.G1.g, e) for e in E["h"]] == P["h"]


def _pts(points):
    xy = np.zeros((len(points), 8), dtype=np.uint64)
    inf = np.zeros(len(points), dtype=np.uint8)
    for i, p in enumerate(points):
        if p is None:
            inf[i] = 1
            xy[i, 4:] = B.int_to_limbs(B.to_mont(1, B.FQ))   # (0, R, inf) as the reference stores the identity
        else:
            xy[i, :4] = B.int_to_limbs(B.to_mont(p[0], B.FQ))
            xy[i, 4:] = B.int_to_limbs(B.to_mont(p[1], B.FQ))
    return xy, inf


def _pts_g2(points):
    xy = np.zeros((len(points), 16), dtype=np.uint64)
    inf = np.zeros(len(points), dtype=np.uint8)
    for i, p in enumerate(points):
        if p is None:
            inf[i] = 1
            xy[i, 8:12] = B.int_to_limbs(B.to_mont(1, B.FQ))
        else:
            for j, v in enumerate((p[0][0], p[0][1], p[1][0], p[1][1])):
                xy[i, 4 * j:4 * j + 4] = B.int_to_limbs(B.to_mont(v, B.FQ))
    return xy, inf


def _enc(aff):
    return np.asarray(aff[:-1], dtype="<u8").tobytes() + bytes([int(aff[-1])])


@pytest.mark.gpu
def test_whole_proof_bytes_identical_with_all_msms_on_gpu(fixture):
    """prover.rs:51-98 with all eight MSMs (six G1, two G2) on the device: Proof {a, b, c} byte for byte."""
    import kogarashi_b200 as k
    from kogarashi_b200.groth16 import Groth16Prover
    k.init()
    P, trap, uvw, proof, internals = fixture
    vk = P["vk"]
    one = lambda p: _pts([p])[0][0]
    one2 = lambda p: _pts_g2([p])[0][0]
    prover = Groth16Prover(one(vk["delta_g1"]), one(vk["alpha_g1"]), one(vk["beta_g1"]), *_pts(P["a"]), *_pts(P["b_g1"]), *_pts(P["h"]), *_pts(P["l"]),
                           one2(vk["delta_g2"]), one2(vk["beta_g2"]), *_pts_g2(P["b_g2"]))
    A, Bp, C = prover.proof(internals["q"], internals["inputs"], internals["aux"], internals["r"], internals["s"])
    assert _enc(Bp) == G.encode_g2(proof["b"])
    assert _enc(A) + _enc(Bp) + _enc(C) == G.proof_bytes(proof)
    # the two G2 MSMs of prover.rs:64-65 one by one (identity CRS entries, the `[l..]` slice)
    l = internals["l"]
    sc = lambda v: np.array([B.int_to_limbs(B.to_mont(x, B.FR)) for x in v], dtype=np.uint64).reshape(-1, 4)
    for pts, coeffs in ((P["b_g2"], internals["inputs"]), (P["b_g2"][l:], internals["aux"])):
        xy, inf = _pts_g2(pts)
        got = k.to_affine(k.BN254_G2, k.msm_curve_addition(xy, sc(coeffs), curve=k.BN254_G2, inf=inf))
        assert _enc(got) == G.encode_g2(G.g2_msm(pts, coeffs))
    prover.free()


@pytest.mark.gpu
def test_proof_bytes_identical_with_g1_msms_on_gpu(fixture):
    import kogarashi_b200 as k
    from kogarashi_b200.groth16 import Groth16G1Prover
    k.init()
    P, trap, uvw, proof, internals = fixture
    vk = P["vk"]
    one = lambda p: _pts([p])[0][0]
    prover = Groth16G1Prover(one(vk["delta_g1"]), one(vk["alpha_g1"]), one(vk["beta_g1"]), *_pts(P["a"]), *_pts(P["b_g1"]), *_pts(P["h"]), *_pts(P["l"]))
    A, C = prover.commitments(internals["q"], internals["inputs"], internals["aux"], internals["r"], internals["s"])

    def enc(aff):
        return np.asarray(aff[:8], dtype="<u8").tobytes() + bytes([int(aff[8])])

    assert enc(A) == G.encode_g1(proof["a"])
    assert enc(C) == G.encode_g1(proof["c"])
    gpu_proof = enc(A) + G.encode_g2(proof["b"]) + enc(C)
    assert gpu_proof == G.proof_bytes(proof)
    # the six MSMs of prover.rs:51-62 one by one through the reference-shaped entry point (3-point MSMs, c = 1 / 3 windows,
    # identity CRS entries, len(bases) > len(coeffs), the `[l..]` slice)
    from oracle import oracle as A_
    l = internals["l"]
    sc = lambda v: np.array([B.int_to_limbs(B.to_mont(x, B.FR)) for x in v], dtype=np.uint64).reshape(-1, 4)
    for pts, coeffs in ((P["h"], internals["q"]), (P["l"], internals["aux"]), (P["a"], internals["inputs"]), (P["a"][l:], internals["aux"]),
                        (P["b_g1"], internals["inputs"]), (P["b_g1"][l:], internals["aux"])):
        xy, inf = _pts(pts)
        got = k.to_affine(k.BN254_G1, k.msm_curve_addition(xy, sc(coeffs), curve=k.BN254_G1, inf=inf))
        exp = G.G1.msm(pts, coeffs)
        assert enc(got) == G.encode_g1(exp)
    prover.free()


@pytest.mark.gpu
@pytest.mark.parametrize("steps,precompute", [(341, False), (1365, True)])
def test_scaled_prover_on_gpu(steps, precompute):
    """The example's function chained to 2^10 / 2^12 constraints.  CRS points come from the device fixed-base
    multiplication of the exponents, H from the device NTT, A, B and C from the device MSMs; all are checked against their
    discrete logs computed from the toxic waste (and the Groth16 equation in the exponent), i.e. against an oracle that
    shares no code with the GPU path."""
    import kogarashi_b200 as k
    from kogarashi_b200 import msm as M
    from kogarashi_b200.groth16 import Groth16Prover
    from oracle import oracle as A_
    k.init()
    cs, out = G.chain_circuit(steps, 3)
    E, trap, uvw = G.crs_exponents(cs, B.XorShift128(SETUP_SEED))
    mont = lambda vals: np.array([B.int_to_limbs(B.to_mont(v, B.FR)) for v in vals], dtype=np.uint64).reshape(-1, 4)

    def points(exps):
        xy = M.fixed_base_mul(k.BN254_G1, mont(exps))           # e * G on the device; exponent 0 -> identity -> (0, 0)
        inf = np.array([1 if e == 0 else 0 for e in exps], dtype=np.uint8)
        return xy, inf

    def points_g2(exps):
        return M.fixed_base_mul(k.BN254_G2, mont(exps)), np.array([1 if e == 0 else 0 for e in exps], dtype=np.uint8)

    vk = points([trap["delta"], trap["alpha"], trap["beta"]])[0]
    vk2 = points_g2([trap["delta"], trap["beta"]])[0]
    prover = Groth16Prover(vk[0], vk[1], vk[2], *points(E["a"]), *points(E["b_g1"]), *points(E["h"]), *points(E["l"]),
                           vk2[0], vk2[1], *points_g2(E["b_g1"]), precompute=precompute)   # b_g2 has the exponents of b_g1 on the G2 generator
    a_ev, b_ev, c_ev = (mont(v) for v in cs.evaluate())
    rng = B.XorShift128(PROVE_SEED)
    r, s = rng.random_field(B.FR), rng.random_field(B.FR)
    A, Bp, C, q = prover.prove_from_evaluations(E["k"], a_ev, b_ev, c_ev, mont(cs.x), mont(cs.w), r, s)
    # H coefficients: device NTT pipeline == restated reference FFT pipeline
    q_ref, n_ref = A_.groth16_h(E["k"], a_ev, b_ev, c_ev)
    assert q.shape[0] == n_ref and (q == q_ref[:n_ref]).all()
    q_int = [B.from_mont(B.limbs_to_int(x), B.FR) for x in q]
    a_exp, b_exp, c_exp, pairing_ok = G.expected_exponents(trap, uvw, cs.x, cs.w, q_int, E["n"], r, s)
    assert pairing_ok

    def enc(aff):
        return np.asarray(aff[:8], dtype="<u8").tobytes() + bytes([int(aff[8])])

    assert enc(A) == G.encode_g1(G.G1.mul(G.G1.g, a_exp))
    assert enc(C) == G.encode_g1(G.G1.mul(G.G1.g, c_exp))
    assert _enc(Bp) == G.encode_g2(G.g2_mul(G.G2_GEN, b_exp))
    # the same call with every lane driven from the calling thread, and the unfused sequence (H, then one batch): same proof
    k.set_param("lane_threads", 0)
    try:
        A0, B0, C0, q0 = prover.prove_from_evaluations(E["k"], a_ev, b_ev, c_ev, mont(cs.x), mont(cs.w), r, s)
    finally:
        k.set_param("lane_threads", 1)
    assert (A0 == A).all() and (B0 == Bp).all() and (C0 == C).all() and (q0 == q).all()
    A1, B1, C1 = prover.prove(q, mont(cs.x), mont(cs.w), r, s)
    assert (A1 == A).all() and (B1 == Bp).all() and (C1 == C).all()
    prover.free()


@pytest.mark.gpu
def test_prover_refuses_identity_delta(fixture):
    """prover.rs:67-69: `if vk.delta_g1.is_identity() || vk.delta_g2.is_identity() { return Err(Error::ProverSubVersionCrsAttack) }` —
    in the reference's encoding (0, R, is_infinity = true), with the flag passed through vk_inf / vk_g2_inf, and as all-zero coordinates."""
    import kogarashi_b200 as k
    from kogarashi_b200.groth16 import Groth16G1Prover, Groth16Prover, ProverSubVersionCrsAttack
    k.init()
    P, trap, uvw, proof, internals = fixture
    vk = P["vk"]
    one = lambda p: _pts([p])[0][0]
    one2 = lambda p: _pts_g2([p])[0][0]
    g1_args = (*_pts(P["a"]), *_pts(P["b_g1"]), *_pts(P["h"]), *_pts(P["l"]))
    ident_g1 = np.zeros(8, dtype=np.uint64)
    ident_g1[4:] = B.int_to_limbs(B.to_mont(1, B.FQ))                     # (0, R): the reference's affine identity coordinates
    with pytest.raises(ProverSubVersionCrsAttack):
        Groth16G1Prover(ident_g1, one(vk["alpha_g1"]), one(vk["beta_g1"]), *g1_args, vk_inf=[1, 0, 0])
    with pytest.raises(ProverSubVersionCrsAttack):
        Groth16G1Prover(np.zeros(8, dtype=np.uint64), one(vk["alpha_g1"]), one(vk["beta_g1"]), *g1_args)
    ident_g2 = np.zeros(16, dtype=np.uint64)
    ident_g2[8:12] = B.int_to_limbs(B.to_mont(1, B.FQ))
    with pytest.raises(ProverSubVersionCrsAttack):
        Groth16Prover(one(vk["delta_g1"]), one(vk["alpha_g1"]), one(vk["beta_g1"]), *g1_args, ident_g2, one2(vk["beta_g2"]), *_pts_g2(P["b_g2"]), vk_g2_inf=[1, 0])
    # an identity alpha / beta is legal for the prover: the flag reaches the device and the proof equals the one computed without that term
    prover = Groth16G1Prover(one(vk["delta_g1"]), ident_g1, one(vk["beta_g1"]), *g1_args, vk_inf=[0, 1, 0])
    a_pt, _ = prover.commitments(internals["q"], internals["inputs"], internals["aux"], internals["r"], internals["s"])
    ref = Groth16G1Prover(one(vk["delta_g1"]), one(vk["alpha_g1"]), one(vk["beta_g1"]), *g1_args)
    a_ref, _ = ref.commitments(internals["q"], internals["inputs"], internals["aux"], internals["r"], internals["s"])
    minus_alpha = np.concatenate([one(vk["alpha_g1"])[:4], A_field_neg(one(vk["alpha_g1"])[4:])])
    exp = k.to_affine(k.BN254_G1, k.msm_curve_addition(np.stack([a_ref[:8], minus_alpha]), np.array([[1, 0, 0, 0], [1, 0, 0, 0]], dtype=np.uint64),
                                                       curve=k.BN254_G1, scalar_fmt=k.SCALARS_CANONICAL))
    assert (a_pt == exp).all()
    prover.free()
    ref.free()


def A_field_neg(y):
    from oracle import oracle as A_
    return A_.field_op(A_.FIELD_FQ, "neg", np.asarray(y, dtype=np.uint64))
