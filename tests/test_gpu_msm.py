"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libkgr_msm.so via kogarashi_b200) and is compared with the CPU oracle / committed goldens as
normalised affine points, bit for bit."""
import os

import numpy as np
import pytest
from conftest import golden_case_names, same_affine

from oracle import oracle as A
from oracle import pyref as B

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def k():
    import kogarashi_b200 as kk
    kk.init()
    yield kk
    kk.set_param("window_bits", 0)
    kk.set_param("chunk", 0)


FIELD_OPS = {"add": 0, "sub": 1, "mul": 2, "square": 3, "neg": 4, "mont_reduce": 5, "to_mont": 6, "invert": 7, "double": 8}


@pytest.mark.parametrize("fid", [A.FIELD_FQ, A.FIELD_FR])
def test_safegcd_inverse_bit_exact(k, fid):
    """modinv.cuh (batched divsteps) against the reference's Fermat inversion (normal.rs:256-287) and against a * a^-1 = 1."""
    from kogarashi_b200 import msm as M
    p = B.FQ if fid == A.FIELD_FQ else B.FR
    n = 1 << 14
    a = A.random_field(fid, n, seed=bytes(range(30, 46)))
    for i, v in enumerate([0, 1, 2, p - 1, p - 2, (1 << 253) % p, 3 << 200]):
        a[i] = B.int_to_limbs(v % p)
    fast = M.test_field_op(fid, 9, a)
    fermat = M.test_field_op(fid, 7, a)
    assert (fast == fermat).all()
    for i in list(range(0, n, 257)) + list(range(8)):
        exp = A.field_op(fid, "invert", a[i])
        assert (fast[i] == (exp if exp is not None else np.zeros(4, dtype=np.uint64))).all(), i


@pytest.mark.parametrize("fid", [A.FIELD_FQ, A.FIELD_FR])
def test_field_ops_bit_exact(k, fid):
    """The PTX carry chains (field.cuh) against zkstd's limb arithmetic on 2^16 random + edge operands."""
    from kogarashi_b200 import msm as M
    p = B.FQ if fid == A.FIELD_FQ else B.FR
    n = 1 << 16
    a = A.random_field(fid, n, seed=bytes(range(16)))
    b = A.random_field(fid, n, seed=bytes(range(1, 17)))
    edge = [0, 1, 2, p - 1, p - 2, (1 << 253) % p, (1 << 254) % p]
    for i, v in enumerate(edge):
        a[i] = B.int_to_limbs(v)
        b[(i * 3) % len(edge)] = B.int_to_limbs(edge[-1 - i])
    for name, op in FIELD_OPS.items():
        m = n if name != "invert" else 2048
        got = M.test_field_op(fid, op, a[:m], b[:m])
        idx = list(range(0, m, 1 if m <= 4096 else 97)) + list(range(len(edge)))
        for i in idx:
            exp = A.field_op(fid, name, a[i], b[i])
            if exp is None:
                exp = np.zeros(4, dtype=np.uint64)  # fp_inv maps 0 -> 0 where the reference returns None
            assert (got[i] == exp).all(), (name, i)


@pytest.mark.parametrize("curve", [A.BN254_G1, A.GRUMPKIN])
def test_point_ops_bit_exact(k, curve):
    from kogarashi_b200 import msm as M
    n = 256
    a = A.random_points(curve, n, seed=bytes(range(16)))
    b = A.random_points(curve, n, seed=bytes(range(2, 18)))
    a_inf = np.zeros(n, dtype=np.uint8)
    b_inf = np.zeros(n, dtype=np.uint8)
    b[0:8] = a[0:8]                                       # P + P -> doubling branch
    for i in range(8, 16):                                # P + (-P) -> identity
        b[i] = a[i]
        b[i, 4:] = A.field_op(A.BASE_FIELD[curve], "neg", a[i, 4:])
    a_inf[16:20] = 1
    b_inf[18:24] = 1
    one = A.field_op(A.BASE_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))

    def proj(pt, inf):
        if inf:
            return np.concatenate([np.zeros(4, dtype=np.uint64), one, np.zeros(4, dtype=np.uint64)])
        return np.concatenate([pt, one])

    got_add = M.test_point_op(curve, 0, a, b, a_inf, b_inf)
    got_dbl = M.test_point_op(curve, 1, a, b, a_inf, b_inf)
    got_a3b = M.test_point_op(curve, 2, a, b, a_inf, b_inf)
    for i in range(n):
        pa, pb = proj(a[i], a_inf[i]), proj(b[i], b_inf[i])
        assert same_affine(A.to_affine(curve, got_add[i]), A.to_affine(curve, A.point_op(curve, 0, pa, pb))), i
        assert same_affine(A.to_affine(curve, got_dbl[i]), A.to_affine(curve, A.point_op(curve, 1, pa))), i
        b3 = A.point_op(curve, 0, A.point_op(curve, 1, pb), pb)
        assert same_affine(A.to_affine(curve, got_a3b[i]), A.to_affine(curve, A.point_op(curve, 0, pa, b3))), i
    # the quad-cooperative addition / doubling of the reduction tail (curve.cuh: four lanes per point operation): general case on non-trivial
    # representatives, the equal-points branch (two representatives of a + b), the opposite-points branch, identity operands
    got_q3 = M.test_point_op(curve, 3, a, b, a_inf, b_inf)
    got_q4 = M.test_point_op(curve, 4, a, b, a_inf, b_inf)
    got_q5 = M.test_point_op(curve, 5, a, b, a_inf, b_inf)
    ident = A.to_affine(curve, proj(a[0], 1))
    for i in range(n):
        pa, pb = proj(a[i], a_inf[i]), proj(b[i], b_inf[i])
        b3 = A.point_op(curve, 0, A.point_op(curve, 1, pb), pb)
        assert same_affine(A.to_affine(curve, got_q3[i]), A.to_affine(curve, A.point_op(curve, 0, pa, b3))), i
        assert same_affine(A.to_affine(curve, got_q4[i]), A.to_affine(curve, A.point_op(curve, 1, A.point_op(curve, 0, pa, pb)))), i
        assert same_affine(A.to_affine(curve, got_q5[i]), ident), i


@pytest.mark.parametrize("name", golden_case_names())
def test_msm_golden_vectors(k, golden, name):
    curve = A.BN254_G1 if name.startswith("g1_") else A.GRUMPKIN
    pts, sc, inf, aff = (golden[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
    got = k.msm_curve_addition(pts, sc, curve=curve, inf=inf)
    assert same_affine(k.to_affine(curve, got), aff)
    assert same_affine(A.to_affine(curve, got), aff)      # the oracle's normalisation agrees with ours
    # canonical-format scalars give the same element
    can = np.stack([A.field_op(A.SCALAR_FIELD[curve], "mont_reduce", s) for s in sc]) if len(sc) else sc
    got2 = k.msm_curve_addition(pts, can, curve=curve, inf=inf, scalar_fmt=k.SCALARS_CANONICAL)
    assert same_affine(k.to_affine(curve, got2), aff)


@pytest.mark.parametrize("c,chunk", [(1, 16), (2, 3), (5, 1), (8, 7), (11, 64), (13, 256)])
def test_msm_golden_under_forced_shapes(k, golden, c, chunk):
    """Every window size / chunk length must give the same group element (signed digits, chunk stitching)."""
    k.set_param("window_bits", c)
    k.set_param("chunk", chunk)
    try:
        for name in ("g1_uniform_1024", "gr_skewed_128", "g1_dup_neg_96", "gr_identity_bases_40", "g1_cancel_32", "g1_rm1_scalars"):
            curve = A.BN254_G1 if name.startswith("g1_") else A.GRUMPKIN
            pts, sc, inf, aff = (golden[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve, inf=inf)), aff), (name, c, chunk)
    finally:
        k.set_param("window_bits", 0)
        k.set_param("chunk", 0)


@pytest.mark.parametrize("param,value", [("final_on_device", 1), ("running_sum_stop", 1), ("running_sum_stop", 1 << 20), ("reduce_fanin", 4), ("sort_mode", 0),
                                         ("sort_mode", 1), ("sort_mode", 2), ("reduce_mode", 0), ("reduce_mode", 1), ("affine_levels", 1), ("affine_levels", 2), ("affine_levels", 4)])
def test_msm_golden_under_reduce_variants(k, golden, param, value):
    """Device-side Horner, pure running-sum reduction, pure weighting-pass reduction, small fan-in: same element."""
    defaults = {"final_on_device": 0, "running_sum_stop": 4096, "reduce_fanin": 16, "sort_mode": -1, "reduce_mode": 1, "affine_levels": -1}
    k.set_param(param, value)
    try:
        for cbits in (0, 10):  # c = 10: 512 buckets per window, enough for the fold reduce to engage
            k.set_param("window_bits", cbits)
            for name in ("g1_uniform_1024", "gr_skewed_128", "g1_dup_neg_96", "gr_cancel_32", "gr_uniform_100"):
                curve = A.BN254_G1 if name.startswith("g1_") else A.GRUMPKIN
                pts, sc, inf, aff = (golden[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
                assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve, inf=inf)), aff), (name, param, value, cbits)
    finally:
        k.set_param(param, defaults[param])
        k.set_param("window_bits", 0)


@pytest.mark.parametrize("curve,logn", [(A.BN254_G1, 10), (A.GRUMPKIN, 12), (A.BN254_G1, 16)])
def test_msm_vs_oracle_seeded(k, curve, logn):
    """Same seeded inputs through the CUDA path and through the restated reference algorithm (config #1 is 2^16)."""
    n = 1 << logn
    pool = A.random_points(curve, min(n, 4096), seed=bytes(range(3, 19)))
    pts = np.tile(pool, (n // pool.shape[0], 1))
    sc = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(4, 20)))
    exp = A.to_affine(curve, A.msm(curve, pts, sc))
    got = k.msm_curve_addition(pts, sc, curve=curve)
    assert same_affine(k.to_affine(curve, got), exp)
    # registered bases + offset window (prover.rs:59: msm(&params.a[l..], aux))
    bases = k.Bases(curve, pts)
    off = 37
    got2 = k.msm_curve_addition(bases, sc[: n - off], base_off=off)
    exp2 = A.to_affine(curve, A.msm(curve, pts[off:], sc[: n - off]))
    assert same_affine(k.to_affine(curve, got2), exp2)
    bases.free()


@pytest.mark.parametrize("window_bits", [0, 3, 7, 12])
def test_precomputed_bases_same_element(k, golden, window_bits):
    """kgr_bases_precompute (window-collapsed mode on the 2^(c*w) * P_i table): same group element for every golden case,
    for sub-ranges (base_off) and for both scalar formats."""
    for name in ("g1_uniform_1024", "gr_uniform_100", "gr_skewed_128", "g1_dup_neg_96", "gr_identity_bases_40", "g1_cancel_32", "g1_rm1_scalars",
                 "g1_zero_scalars", "gr_uniform_1", "g1_uniform_3"):
        curve = A.BN254_G1 if name.startswith("g1_") else A.GRUMPKIN
        pts, sc, inf, aff = (golden[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
        bases = k.Bases(curve, pts, inf).precompute(window_bits)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), aff), (name, window_bits)
        if len(sc) > 8:
            off = 5
            exp = A.to_affine(curve, A.msm(curve, pts[off:], sc[: len(sc) - off], inf=inf[off:]))
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc[: len(sc) - off], base_off=off)), exp), (name, "off")
        bases.free()


def test_precomputed_bases_checksum_2p18(k):
    curve, n = A.BN254_G1, 1 << 18
    cm = B.CURVES[curve]
    bases, ks = k.Bases.generate(curve, n, seed=21, return_scalars=True)
    bases.precompute()
    sc = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(40, 56)))
    got = k.msm_curve_addition(bases, sc)
    g = A.generator(curve)
    one = A.field_op(A.BASE_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    s = _dot_mod(ks, sc, cm.r)
    exp = A.to_affine(curve, A.scalar_point(curve, np.concatenate([g, one]), np.array(B.int_to_limbs(B.to_mont(s, cm.r)), dtype=np.uint64)))
    assert same_affine(k.to_affine(curve, got), exp)
    bases.free()


def _dot_mod(ks, sc, r):
    tot = 0
    for a, b in zip(ks, sc):
        tot += B.from_mont(B.limbs_to_int(a), r) * B.from_mont(B.limbs_to_int(b), r)
    return tot % r


@pytest.mark.parametrize("curve,logn", [(A.BN254_G1, 20), (A.GRUMPKIN, 20)])
def test_msm_full_size_checksum(k, curve, logn):
    """BASELINE configs #2/#3 (2^20): bases are k_i*G with known k_i, so the MSM must equal (sum k_i s_i) * G —
    a size-independent check against one oracle scalar multiplication."""
    cm = B.CURVES[curve]
    n = 1 << logn
    bases, ks = k.Bases.generate(curve, n, seed=11, return_scalars=True)
    # spot-check generated bases against the oracle: P_i == k_i * G
    g = A.generator(curve)
    one = A.field_op(A.BASE_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    gp = np.concatenate([g, one])
    sc = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(5, 21)))
    got = k.msm_curve_addition(bases, sc)
    s = _dot_mod(ks, sc, cm.r)
    exp = A.to_affine(curve, A.scalar_point(curve, gp, np.array(B.int_to_limbs(B.to_mont(s, cm.r)), dtype=np.uint64)))
    assert same_affine(k.to_affine(curve, got), exp)
    # linearity at full size: msm(P, a) + msm(P, b) == msm(P, a + b)
    sc2 = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(6, 22)))
    got_b = k.msm_curve_addition(bases, sc2)
    s2 = (s + _dot_mod(ks, sc2, cm.r)) % cm.r
    exp_sum = A.to_affine(curve, A.scalar_point(curve, gp, np.array(B.int_to_limbs(B.to_mont(s2, cm.r)), dtype=np.uint64)))
    assert same_affine(k.to_affine(curve, k.proj_add(curve, got, got_b)), exp_sum)
    bases.free()


def test_fixed_base_mul_matches_oracle(k):
    from kogarashi_b200 import msm as M
    for curve in (A.BN254_G1, A.GRUMPKIN):
        pts, ks = A.random_points(curve, 64, seed=bytes(range(9, 25)), return_scalars=True)
        got = M.fixed_base_mul(curve, ks)
        assert (got == pts).all()


@pytest.mark.parametrize("curve", [A.BN254_G1, A.GRUMPKIN])
def test_pedersen_commit_matches_reference_fold(k, curve):
    """nova/src/pedersen.rs:15-20: commit == fold of sum + g_i * m_i (oracle restatement), zip semantics."""
    g = A.random_points(curve, 129, seed=bytes(range(7, 23)))          # 2^7 + 1 generators
    m = A.random_field(A.SCALAR_FIELD[curve], 100, seed=bytes(range(8, 24)))
    m[::2] = 0                                                          # witness-like: many zeros / ones
    m[1::4] = A.field_op(A.SCALAR_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    ck = k.PedersenCommitment(curve, g)
    assert same_affine(ck.commit(m), A.pedersen_commit(curve, g, m))
    longer = A.random_field(A.SCALAR_FIELD[curve], 200, seed=bytes(range(10, 26)))
    assert same_affine(ck.commit(longer), A.pedersen_commit(curve, g, longer))


def test_skewed_scalars_large(k):
    """Nova-shaped scalars (half zeros, a quarter ones) at 2^15 — one bucket holds a quarter of the points."""
    curve, n = A.GRUMPKIN, 1 << 15
    cm = B.CURVES[curve]
    bases, ks = k.Bases.generate(curve, n, seed=5, return_scalars=True)
    sc = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(12, 28)))
    sc[::2] = 0
    sc[1::4] = A.field_op(A.SCALAR_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    got = k.msm_curve_addition(bases, sc)
    g = A.generator(curve)
    one = A.field_op(A.BASE_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    s = _dot_mod(ks, sc, cm.r)
    exp = A.to_affine(curve, A.scalar_point(curve, np.concatenate([g, one]), np.array(B.int_to_limbs(B.to_mont(s, cm.r)), dtype=np.uint64)))
    assert same_affine(k.to_affine(curve, got), exp)


def test_abi_error_codes(k, golden):
    """Error behaviour at the C boundary: negative KGR_E_* codes with a message, no exception, engine usable afterwards."""
    import ctypes
    from kogarashi_b200 import _lib
    L = _lib.lib()
    u64p = ctypes.POINTER(ctypes.c_uint64)
    pts, sc, aff = golden["g1_uniform_32_pts"], golden["g1_uniform_32_sc"], golden["g1_uniform_32_aff"]
    out = np.zeros(12, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(u64p)
    bases = k.Bases(A.BN254_G1, pts)
    E_ARG = -3
    assert L.kgr_msm(bases._h, 30, p(sc), 0, 5, p(out)) == E_ARG and b"range" in L.kgr_last_error()        # off + n beyond the vector
    assert L.kgr_msm(bases._h, 0, p(sc), 7, 32, p(out)) == E_ARG and b"scalar format" in L.kgr_last_error()
    assert L.kgr_msm(bases._h, 0, None, 0, 32, p(out)) == E_ARG
    assert L.kgr_msm(None, 0, p(sc), 0, 32, p(out)) == E_ARG
    assert L.kgr_msm_oneshot(9, p(pts), None, 32, p(sc), 0, 32, p(out)) == E_ARG and b"curve" in L.kgr_last_error()
    assert L.kgr_set_param(b"no_such_knob", 1) == E_ARG
    assert L.kgr_set_param(b"reduce_fanin", 3) == E_ARG
    assert L.kgr_bases_precompute(bases._h, 99) == E_ARG
    h = ctypes.c_void_p()
    assert L.kgr_bases_register(5, p(pts), None, 32, ctypes.byref(h)) != 0
    # the engine still works and an empty MSM is the identity (msm.rs:45-47 folds nothing)
    assert same_affine(k.to_affine(A.BN254_G1, k.msm_curve_addition(bases, sc)), aff)
    assert L.kgr_msm(bases._h, 32, p(sc), 0, 0, p(out)) == 0
    assert int(k.to_affine(A.BN254_G1, out)[8]) == 1
    bases.free()


@pytest.mark.parametrize("curve", [A.BN254_G1, A.BN254_G2])
def test_oneshot_pieces_same_element(k, curve):
    """Large kgr_msm_oneshot calls on one device are cut into pipelined pieces ("oneshot_split"): same group element as the
    registered-bases call for 1, 2, 3 and 4 pieces, with identity bases and a ragged length."""
    n = (1 << 18) + 12345
    bases = k.Bases.generate(curve, n, seed=31)
    pts = bases.download()
    inf = np.zeros(n, dtype=np.uint8)
    inf[[0, 5, n // 2, n - 1]] = 1
    pts_inf = pts.copy()
    pts_inf[inf == 1] = 0
    sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(9, 25)))
    ref = k.to_affine(curve, k.msm_curve_addition(bases, sc))
    reg_inf = k.Bases(curve, pts_inf, inf)
    ref_inf = k.to_affine(curve, k.msm_curve_addition(reg_inf, sc))
    try:
        for pieces in (1, 2, 3, 4, 0):
            k.set_param("oneshot_split", pieces)
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), ref), pieces
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts_inf, sc, curve=curve, inf=inf)), ref_inf), pieces
            # registered bases, host scalars: the same cut (scalar uploads under the pipeline), also with an offset window
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(reg_inf, sc)), ref_inf), pieces
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc[: n - 77], base_off=77)),
                               k.to_affine(curve, k.msm_curve_addition(pts[77:], sc[: n - 77], curve=curve))), pieces
        if curve == A.BN254_G1:   # window-collapsed table: pieces index the same table at shifted columns
            bases.precompute(0)
            for pieces in (1, 3):
                k.set_param("oneshot_split", pieces)
                assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), ref), ("table", pieces)
                assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc[: n - 77], base_off=77)),
                                   k.to_affine(curve, k.msm_curve_addition(pts[77:], sc[: n - 77], curve=curve))), ("table", pieces)
    finally:
        k.set_param("oneshot_split", 0)
    bases.free()
    reg_inf.free()


def test_geometric_pieces_tiny_and_lopsided(k):
    """The pieces of a streamed host-buffer call grow geometrically ("oneshot_growth", percent): lopsided cuts, pieces of a single pair, cuts
    that would be empty (skipped) — always the oracle's group element."""
    curve = A.BN254_G1
    try:
        for n in (1, 2, 5, 64, 1000):
            pts = A.random_points(curve, n, seed=bytes(range(3, 19)))
            sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(5, 21)))
            exp = A.to_affine(curve, A.msm(curve, pts, sc))
            bases = k.Bases(curve, pts)
            for pieces in (2, 3, 5, 8):
                for growth in (100, 170, 300, 1000):
                    k.set_param("oneshot_split", pieces)
                    k.set_param("oneshot_growth", growth)
                    assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp), (n, pieces, growth)
                    assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), exp), (n, pieces, growth, "registered")
            bases.free()
    finally:
        k.set_param("oneshot_split", 0)
        k.set_param("oneshot_growth", 0)


@pytest.mark.parametrize("seed", range(16))
def test_random_shapes_vs_oracle(k, seed):
    """Randomised configurations: curve, size, forced window / chunk, sort and reduce variants, scalar mix (zeros, ones, r - 1, uniform),
    duplicated / opposite / identity bases, ragged lengths — each compared with the restated reference algorithm."""
    rng = np.random.default_rng(1000 + seed)
    curve = int(rng.integers(0, 3))
    cl = 8 if curve == A.BN254_G2 else 4
    n = int(rng.choice([1, 2, 3, 7, 33, 100, 257, 1000, 3000]))
    pool = A.random_points(curve, min(n, 64), seed=bytes((seed + i) % 256 for i in range(16)))
    pts = pool[rng.integers(0, pool.shape[0], size=n)]                 # repeated bases
    fid = A.SCALAR_FIELD[curve]
    sc = A.random_field(fid, n, seed=bytes((2 * seed + i) % 256 for i in range(16)))
    one = A.field_op(fid, "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    rm1 = A.field_op(fid, "neg", one)
    kind = rng.integers(0, 4, size=n)
    sc[kind == 0] = 0
    sc[kind == 1] = one
    if seed % 3 == 0:
        sc[kind == 2] = rm1
    inf = (rng.integers(0, 9, size=n) == 0).astype(np.uint8)
    neg = rng.integers(0, 5, size=n) == 0                               # opposite points
    bf = A.FIELD_FR if curve == A.GRUMPKIN else A.FIELD_FQ
    for i in np.nonzero(neg)[0]:
        for h in range(cl, 2 * cl, 4):
            pts[i, h:h + 4] = A.field_op(bf, "neg", pts[i, h:h + 4])
    m = n if seed % 4 else max(1, n - 3)                                # fewer scalars than bases
    params = {"window_bits": int(rng.integers(0, 15)), "chunk": int(rng.choice([0, 1, 2, 5, 16, 64])), "sort_mode": int(rng.integers(-1, 3)),
              "reduce_mode": int(rng.integers(0, 2)), "final_on_device": int(rng.integers(0, 2)), "affine_levels": int(rng.integers(-1, 5))}
    defaults = {"window_bits": 0, "chunk": 0, "sort_mode": -1, "reduce_mode": 1, "final_on_device": 0, "affine_levels": -1}
    exp = A.to_affine(curve, A.msm(curve, pts, sc[:m], inf=inf))
    try:
        for name, v in params.items():
            k.set_param(name, v)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc[:m], curve=curve, inf=inf)), exp), params
        bases = k.Bases(curve, pts, inf)
        if seed % 2:
            bases.precompute(int(rng.integers(1, 12)))
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc[:m])), exp), params
        bases.free()
    finally:
        for name, v in defaults.items():
            k.set_param(name, v)


@pytest.mark.parametrize("c", [2, 3, 7, 9, 12, 14, 16, 18])
def test_radix_partition_sort_forced_window_sizes(k, golden, c):
    """sort_mode = 2 (two block-local radix partitions, kernels_sort.cu) for every split of the bucket index into coarse bins and fine bits:
    the goldens with duplicates / opposite points / identity bases / r-1 scalars, both scalar formats, and the window-table mode."""
    k.set_param("sort_mode", 2)
    k.set_param("window_bits", c)
    try:
        for name in ("g1_uniform_1024", "gr_skewed_128", "g1_dup_neg_96", "gr_identity_bases_40", "g1_cancel_32", "g1_rm1_scalars", "gr_uniform_100"):
            curve = A.BN254_G1 if name.startswith("g1_") else A.GRUMPKIN
            pts, sc, inf, aff = (golden[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve, inf=inf)), aff), (name, c)
            can = np.stack([A.field_op(A.SCALAR_FIELD[curve], "mont_reduce", x) for x in sc])
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, can, curve=curve, inf=inf, scalar_fmt=k.SCALARS_CANONICAL)), aff), (name, c)
            if c <= 12:
                bases = k.Bases(curve, pts, inf).precompute(c)
                assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), aff), (name, c, "table")
                bases.free()
    finally:
        k.set_param("sort_mode", -1)
        k.set_param("window_bits", 0)


@pytest.mark.parametrize("curve,logn", [(A.BN254_G1, 16), (A.GRUMPKIN, 14)])
def test_radix_partition_sort_vs_oracle(k, curve, logn):
    """Seeded inputs with hot buckets (half the scalars zero, a quarter equal) through sort_mode 2 against the restated reference algorithm:
    bins that overflow the shared-memory buffer take the direct-placement path."""
    n = 1 << logn
    rng = np.random.default_rng(99 + logn)
    pool = A.random_points(curve, 2048, seed=bytes(range(9, 25)))
    pts = pool[rng.integers(0, pool.shape[0], size=n)]
    fid = A.SCALAR_FIELD[curve]
    sc = A.random_field(fid, n, seed=bytes(range(40, 56)))
    kind = rng.integers(0, 4, size=n)
    sc[kind == 0] = 0
    sc[kind == 1] = sc[0]
    exp = A.to_affine(curve, A.msm(curve, pts, sc))
    k.set_param("sort_mode", 2)
    try:
        for c in (0, 8, 15):
            k.set_param("window_bits", c)
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp), c
    finally:
        k.set_param("sort_mode", -1)
        k.set_param("window_bits", 0)


@pytest.mark.parametrize("sort_mode", [0, 1, 2])
def test_unreduced_scalars_are_read_modulo_the_modulus(k, sort_mode):
    """Canonical scalars >= r (which to_raw_bytes never produces) and Montgomery limbs that are not fully reduced give the MSM of the
    scalars mod r — not a digit string that silently lost its top bits."""
    curve, n = A.BN254_G1, 64
    rng = np.random.default_rng(5)
    pts = A.random_points(curve, n, seed=bytes(range(16)))
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    raw[0] = np.uint64(0xFFFFFFFFFFFFFFFF)                           # 2^256 - 1
    raw[1] = B.int_to_limbs(B.FR)                                    # exactly r -> 0
    raw[2] = B.int_to_limbs(B.FR + 5)
    vals = [B.limbs_to_int(x) % B.FR for x in raw]
    k.set_param("sort_mode", sort_mode)
    try:
        # canonical: value = raw mod r
        red = np.array([B.int_to_limbs(B.to_mont(v, B.FR)) for v in vals], dtype=np.uint64)
        exp = A.to_affine(curve, A.msm(curve, pts, red))
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, raw, curve=curve, scalar_fmt=k.SCALARS_CANONICAL)), exp)
        # Montgomery: value = raw * R^-1 mod r
        red_m = np.array([B.int_to_limbs(v) for v in vals], dtype=np.uint64)   # limbs of (raw mod r) are a valid Montgomery residue of the same class
        exp_m = A.to_affine(curve, A.msm(curve, pts, red_m))
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, raw, curve=curve)), exp_m)
    finally:
        k.set_param("sort_mode", -1)


def test_gpu_matches_alt_bn128_kat(k):
    """External known-answer vectors (EIP-196 precompile tests, tests/golden/alt_bn128_kat.json) through the C ABI: the XYZZ point kernels,
    one- and two-point MSMs in both scalar formats, a scalar above the group order, and the registered-bases path."""
    from conftest import alt_bn128_kat
    from kogarashi_b200 import msm as M
    _, adds, muls = alt_bn128_kat()
    curve = A.BN254_G1
    z1 = np.zeros(1, dtype=np.uint64)
    one_fr = A.field_op(A.FIELD_FR, "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    for name, a, b, s in adds:
        exp = np.concatenate([s, z1])
        got = M.test_point_op(curve, 0, a.reshape(1, 8), b.reshape(1, 8))
        assert same_affine(k.to_affine(curve, got[0]), exp), name
        pts, ones = np.stack([a, b]), np.stack([one_fr, one_fr])
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, ones, curve=curve)), exp), name
        can = np.array([[1, 0, 0, 0], [1, 0, 0, 0]], dtype=np.uint64)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, can, curve=curve, scalar_fmt=k.SCALARS_CANONICAL)), exp), name
    for name, p, kk, prod in muls:
        exp = np.concatenate([prod, z1])
        km = np.array(B.int_to_limbs(B.to_mont(kk % B.FR, B.FR)), dtype=np.uint64).reshape(1, 4)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(p.reshape(1, 8), km, curve=curve)), exp), name
        raw = np.array(B.int_to_limbs(kk), dtype=np.uint64).reshape(1, 4)     # canonical bytes as given, even when >= r (cdetrio1: 2^256 - 1)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(p.reshape(1, 8), raw, curve=curve, scalar_fmt=k.SCALARS_CANONICAL)), exp), name
        bases = k.Bases(curve, np.tile(p, (300, 1)))                           # k * P = sum of 300 copies with scalars that add up to k
        parts = [int(x) for x in np.random.default_rng(3).integers(0, 1 << 62, size=299)]
        parts.append((kk - sum(parts)) % B.FR)
        sc = np.array([B.int_to_limbs(B.to_mont(v, B.FR)) for v in parts], dtype=np.uint64)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), exp), name
        bases.free()


def test_gpu_matches_reference_binary_vectors(k):
    """The CUDA path against vectors printed by the Rust reference itself (skipped until tests/golden/msm_vectors_ref.npz exists)."""
    from test_oracle import _ref_vectors
    z, names = _ref_vectors()
    for name in names:
        got = k.msm_curve_addition(z[name + "_pts"], z[name + "_sc"], curve=A.BN254_G1, inf=z[name + "_inf"])
        assert same_affine(k.to_affine(A.BN254_G1, got), z[name + "_aff"]), name


@pytest.mark.parametrize("levels", [1, 2, 3, 5])
def test_batched_affine_levels_special_cases(k, golden, levels):
    """affine_levels = r (affine_kernels.cuh): pairs inside a bucket that are equal (tangent), opposite (identity), identities themselves, odd
    bucket lengths, empty buckets, a single bucket with everything in it — every golden, small windows so that buckets are long."""
    k.set_param("affine_levels", levels)
    k.set_param("dense_x", levels & 1)   # odd levels: the level-0 denominators read the dense copy of the x coordinates (automatic only above 2^20 points)
    try:
        for c in (2, 5, 0):
            k.set_param("window_bits", c)
            for name in ("g1_uniform_1024", "gr_skewed_128", "g1_dup_neg_96", "gr_identity_bases_40", "g1_cancel_32", "g1_rm1_scalars", "gr_uniform_100", "gr_cancel_32"):
                curve = A.BN254_G1 if name.startswith("g1_") else A.GRUMPKIN
                pts, sc, inf, aff = (golden[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
                assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve, inf=inf)), aff), (name, levels, c)
        # all scalars equal: one bucket per window holds every point (the same point 300 times: every pair is a doubling)
        curve = A.BN254_G1
        p = A.random_points(curve, 1, seed=bytes(range(16)))
        pts = np.tile(p, (300, 1))
        sc = np.tile(A.random_field(A.FIELD_FR, 1, seed=bytes(range(1, 17))), (300, 1))
        exp = A.to_affine(curve, A.msm(curve, pts, sc))
        for c in (3, 9):
            k.set_param("window_bits", c)
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp), (levels, c)
        bases = k.Bases(curve, pts).precompute(4)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), exp)
        bases.free()
    finally:
        k.set_param("affine_levels", -1)
        k.set_param("window_bits", 0)
        k.set_param("dense_x", -1)


@pytest.mark.parametrize("curve,logn", [(A.BN254_G1, 16), (A.GRUMPKIN, 15)])
def test_batched_affine_levels_vs_oracle(k, curve, logn):
    """Seeded inputs with repeated bases and hot buckets through 1..4 affine levels against the restated reference algorithm."""
    n = 1 << logn
    rng = np.random.default_rng(7 + logn)
    pool = A.random_points(curve, 512, seed=bytes(range(11, 27)))          # few distinct bases: equal / opposite pairs inside buckets are common
    pts = pool[rng.integers(0, pool.shape[0], size=n)]
    fid = A.SCALAR_FIELD[curve]
    sc = A.random_field(fid, n, seed=bytes(range(60, 76)))
    kind = rng.integers(0, 4, size=n)
    sc[kind == 0] = sc[1]
    exp = A.to_affine(curve, A.msm(curve, pts, sc))
    try:
        for levels, c in ((1, 0), (2, 8), (3, 10), (4, 6)):
            k.set_param("affine_levels", levels)
            k.set_param("window_bits", c)
            k.set_param("dense_x", levels & 1)
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp), (levels, c)
            k.set_param("oneshot_split", 2)   # two streamed pieces: the copy is made per piece
            assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp), (levels, c, "pieces")
            k.set_param("oneshot_split", 0)
    finally:
        k.set_param("affine_levels", -1)
        k.set_param("window_bits", 0)
        k.set_param("dense_x", -1)
        k.set_param("oneshot_split", 0)


def test_handles_die_with_their_kgr_init_and_pinned_arena(k):
    """a handle created before a second kgr_init is refused (its shards name engines that no longer exist) but can still be freed;
    kgr_host_alloc gives page-locked memory that the calls read directly (what the Rust shim's PinnedArena marshals into)."""
    import ctypes

    from kogarashi_b200 import _lib
    curve = A.BN254_G1
    pts = A.random_points(curve, 64, seed=bytes(range(16)))
    sc = A.random_field(A.FIELD_FR, 64, seed=bytes(range(1, 17)))
    exp = A.to_affine(curve, A.msm(curve, pts, sc))
    old = k.Bases(curve, pts)
    assert same_affine(k.to_affine(curve, k.msm_curve_addition(old, sc)), exp)
    k.init([0])                                            # re-numbers the engines: `old` belongs to the previous generation
    with pytest.raises(k.KgrError, match="before the last kgr_init"):
        k.msm_curve_addition(old, sc)
    with pytest.raises(k.KgrError, match="before the last kgr_init"):
        old.precompute(4)
    old.free()                                             # its device memory is released on the device it was allocated on
    fresh = k.Bases(curve, pts)
    assert same_affine(k.to_affine(curve, k.msm_curve_addition(fresh, sc)), exp)
    fresh.free()
    # pinned arena
    L = _lib.lib()
    p_xy, p_sc = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.kgr_host_alloc(pts.nbytes, ctypes.byref(p_xy)))
    _lib.check(L.kgr_host_alloc(sc.nbytes, ctypes.byref(p_sc)))
    ctypes.memmove(p_xy, pts.ctypes.data, pts.nbytes)
    ctypes.memmove(p_sc, sc.ctypes.data, sc.nbytes)
    got = k.msm_oneshot_ptr(curve, p_xy.value, 64, p_sc.value, 64)
    assert same_affine(k.to_affine(curve, got), exp)
    _lib.check(L.kgr_host_free(p_xy))
    _lib.check(L.kgr_host_free(p_sc))


def test_set_param_is_per_thread(k):
    """kgr_set_param changes the calling thread's tuning state only — a window size forced on another thread does not
    reach an MSM issued from this one (the engine's last shape shows the automatic choice), and vice versa."""
    import threading
    curve = A.BN254_G1
    pts = A.random_points(curve, 256, seed=bytes(range(16)))
    sc = A.random_field(A.FIELD_FR, 256, seed=bytes(range(1, 17)))
    exp = A.to_affine(curve, A.msm(curve, pts, sc))
    k.msm_curve_addition(pts, sc, curve=curve)
    auto_c = k.last_timing(0)[1]["c"]
    seen = {}

    def other():
        k.set_param("window_bits", 3)
        seen["aff"] = k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve))
        seen["c"] = k.last_timing(0)[1]["c"]

    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert seen["c"] == 3 and same_affine(seen["aff"], exp)
    assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp)
    assert k.last_timing(0)[1]["c"] == auto_c              # this thread never asked for c = 3
