"""GPU parity tests for the BN254 G2 MSM (next row N3: the b_g2 queries of groth16/src/prover.rs:64-65).  Same C ABI and
pipeline as G1 with Fq2 coordinates; compared with the oracle / committed goldens as normalised affine points, bit for bit."""
import numpy as np
import pytest
from conftest import golden_g2_case_names, same_affine

from oracle import oracle as A
from oracle import pyref as B

pytestmark = pytest.mark.gpu
C2 = A.BN254_G2
R = B.FR


@pytest.fixture(scope="module")
def k():
    import kogarashi_b200 as kk
    kk.init()
    yield kk
    for name in ("window_bits", "chunk"):
        kk.set_param(name, 0)


def _neg(p):
    q = p.copy()
    q[8:12] = A.field_op(A.FIELD_FQ, "neg", p[8:12])
    q[12:16] = A.field_op(A.FIELD_FQ, "neg", p[12:16])
    return q


def _proj(pt, inf):
    one = A.field_op(A.FIELD_FQ, "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    z4 = np.zeros(4, dtype=np.uint64)
    if inf:
        return np.concatenate([z4, z4, one, z4, z4, z4])
    return np.concatenate([pt, one, z4])


def _dot_mod(ks, sc):
    return sum(B.from_mont(B.limbs_to_int(a), R) * B.from_mont(B.limbs_to_int(b), R) for a, b in zip(ks, sc)) % R


def _k_times_g(s):
    return A.to_affine(C2, A.scalar_point(C2, _proj(A.generator(C2), 0), np.array(B.int_to_limbs(B.to_mont(s, R)), dtype=np.uint64)))


def test_point_ops_bit_exact(k):
    """Fq2 arithmetic + XYZZ group law over Fq2 on the device against the reference's projective formulas (oracle)."""
    from kogarashi_b200 import msm as M
    n = 96
    a = A.random_points(C2, n, seed=bytes(range(16)))
    b = A.random_points(C2, n, seed=bytes(range(2, 18)))
    a_inf, b_inf = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    b[0:8] = a[0:8]                                       # P + P -> doubling branch
    for i in range(8, 16):                                # P + (-P) -> identity
        b[i] = _neg(a[i])
    a_inf[16:20] = 1
    b_inf[18:24] = 1
    got_add = M.test_point_op(C2, 0, a, b, a_inf, b_inf)
    got_dbl = M.test_point_op(C2, 1, a, b, a_inf, b_inf)
    got_a3b = M.test_point_op(C2, 2, a, b, a_inf, b_inf)
    for i in range(n):
        pa, pb = _proj(a[i], a_inf[i]), _proj(b[i], b_inf[i])
        assert same_affine(A.to_affine(C2, got_add[i]), A.to_affine(C2, A.point_op(C2, 0, pa, pb))), i
        assert same_affine(A.to_affine(C2, got_dbl[i]), A.to_affine(C2, A.point_op(C2, 1, pa))), i
        b3 = A.point_op(C2, 0, A.point_op(C2, 1, pb), pb)
        assert same_affine(A.to_affine(C2, got_a3b[i]), A.to_affine(C2, A.point_op(C2, 0, pa, b3))), i
        # host-side helpers of the library agree with the oracle as well
        assert same_affine(k.to_affine(C2, got_add[i]), A.to_affine(C2, got_add[i])), i
        assert same_affine(k.to_affine(C2, k.proj_add(C2, pa, pb)), A.to_affine(C2, A.point_op(C2, 0, pa, pb))), i
    # quad-cooperative addition / doubling (reduction tail): general case, equal-points and opposite-points branches, identity operands
    got_q3 = M.test_point_op(C2, 3, a, b, a_inf, b_inf)
    got_q4 = M.test_point_op(C2, 4, a, b, a_inf, b_inf)
    got_q5 = M.test_point_op(C2, 5, a, b, a_inf, b_inf)
    ident = A.to_affine(C2, _proj(a[0], 1))
    for i in range(n):
        pa, pb = _proj(a[i], a_inf[i]), _proj(b[i], b_inf[i])
        b3 = A.point_op(C2, 0, A.point_op(C2, 1, pb), pb)
        assert same_affine(A.to_affine(C2, got_q3[i]), A.to_affine(C2, A.point_op(C2, 0, pa, b3))), i
        assert same_affine(A.to_affine(C2, got_q4[i]), A.to_affine(C2, A.point_op(C2, 1, A.point_op(C2, 0, pa, pb)))), i
        assert same_affine(A.to_affine(C2, got_q5[i]), ident), i


@pytest.mark.parametrize("name", golden_g2_case_names())
def test_msm_golden_vectors(k, golden_g2, name):
    pts, sc, inf, aff = (golden_g2[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
    assert same_affine(k.to_affine(C2, k.msm_curve_addition(pts, sc, curve=C2, inf=inf)), aff)
    bases = k.Bases(C2, pts, inf)
    assert same_affine(k.to_affine(C2, k.msm_curve_addition(bases, sc)), aff)
    canon = np.stack([A.field_op(A.FIELD_FR, "mont_reduce", s) for s in sc]) if len(sc) else sc
    assert same_affine(k.to_affine(C2, k.msm_curve_addition(bases, canon, scalar_fmt=k.SCALARS_CANONICAL)), aff)
    bases.free()


@pytest.mark.parametrize("c,chunk", [(1, 16), (3, 2), (6, 1), (9, 64), (13, 256)])
def test_msm_golden_under_forced_shapes(k, golden_g2, c, chunk):
    k.set_param("window_bits", c)
    k.set_param("chunk", chunk)
    try:
        for name in ("g2_uniform_128", "g2_dup_neg_48", "g2_identity_bases_24", "g2_skewed_64", "g2_cancel_16", "g2_uniform_1"):
            pts, sc, inf, aff = (golden_g2[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
            assert same_affine(k.to_affine(C2, k.msm_curve_addition(pts, sc, curve=C2, inf=inf)), aff), name
    finally:
        k.set_param("window_bits", 0)
        k.set_param("chunk", 0)


@pytest.mark.parametrize("param,value", [("final_on_device", 1), ("running_sum_stop", 1), ("running_sum_stop", 1 << 20), ("reduce_fanin", 4), ("sort_mode", 1),
                                         ("reduce_mode", 0)])
def test_msm_under_reduce_variants(k, param, value):
    defaults = {"final_on_device": 0, "running_sum_stop": 4096, "reduce_fanin": 16, "sort_mode": -1, "reduce_mode": 1}
    n = 1 << 12
    pts = np.tile(A.random_points(C2, 512, seed=bytes(range(3, 19))), (n // 512, 1))
    sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(4, 20)))
    exp = A.to_affine(C2, A.msm(C2, pts, sc))
    k.set_param(param, value)
    try:
        assert same_affine(k.to_affine(C2, k.msm_curve_addition(pts, sc, curve=C2)), exp)
    finally:
        k.set_param(param, defaults[param])


@pytest.mark.parametrize("logn", [10, 14])
def test_msm_vs_oracle_seeded(k, logn):
    n = 1 << logn
    pool = A.random_points(C2, min(n, 1024), seed=bytes(range(3, 19)))
    pts = np.tile(pool, (n // pool.shape[0], 1))
    sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(4, 20)))
    assert same_affine(k.to_affine(C2, k.msm_curve_addition(pts, sc, curve=C2)), A.to_affine(C2, A.msm(C2, pts, sc)))
    bases = k.Bases(C2, pts)
    off = 37   # prover.rs:65: msm(&params.b_g2[l..], aux)
    got = k.msm_curve_addition(bases, sc[: n - off], base_off=off)
    assert same_affine(k.to_affine(C2, got), A.to_affine(C2, A.msm(C2, pts[off:], sc[: n - off])))
    assert (bases.download(5, 7) == pts[5:12]).all()
    bases.free()


@pytest.mark.parametrize("window_bits", [0, 5, 11])
def test_precomputed_bases_same_element(k, golden_g2, window_bits):
    for name in ("g2_uniform_128", "g2_dup_neg_48", "g2_identity_bases_24", "g2_cancel_16", "g2_rm1_scalars", "g2_uniform_1"):
        pts, sc, inf, aff = (golden_g2[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
        bases = k.Bases(C2, pts, inf).precompute(window_bits)
        assert same_affine(k.to_affine(C2, k.msm_curve_addition(bases, sc)), aff), (name, window_bits)
        bases.free()


def test_fixed_base_mul_matches_oracle(k):
    from kogarashi_b200 import msm as M
    pts, ks = A.random_points(C2, 48, seed=bytes(range(9, 25)), return_scalars=True)
    assert (M.fixed_base_mul(C2, ks) == pts).all()


def test_msm_checksum_2p18(k):
    """Bases k_i * G2 generated on the device with known k_i: the MSM must equal (sum k_i s_i) * G2, checked against one oracle
    scalar multiplication; linearity msm(a) + msm(b) == msm(a + b) at the same size; precomputed table gives the same element."""
    n = 1 << 18
    bases, ks = k.Bases.generate(C2, n, seed=11, return_scalars=True)
    spot = bases.download(1000, 4)
    for i in range(4):
        assert same_affine(np.concatenate([spot[i], np.zeros(1, np.uint64)]), _k_times_g(B.from_mont(B.limbs_to_int(ks[1000 + i]), R)))
    sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(5, 21)))
    got = k.msm_curve_addition(bases, sc)
    s = _dot_mod(ks, sc)
    assert same_affine(k.to_affine(C2, got), _k_times_g(s))
    sc2 = A.random_field(A.FIELD_FR, n, seed=bytes(range(6, 22)))
    got_b = k.msm_curve_addition(bases, sc2)
    s2 = (s + _dot_mod(ks, sc2)) % R
    assert same_affine(k.to_affine(C2, k.proj_add(C2, got, got_b)), _k_times_g(s2))
    bases.precompute()
    assert same_affine(k.to_affine(C2, k.msm_curve_addition(bases, sc)), _k_times_g(s))
    bases.free()


def test_msm_batch_equals_sequence(k, golden_g2, golden):
    """kgr_msm_batch: mixed-curve jobs (G1, Grumpkin, G2, empty, offset, precomputed, canonical scalars) overlapped on the lanes of one
    device give exactly the results of the same calls made one by one; more jobs than lanes reuse lanes in order."""
    specs = [("g1_uniform_1024", 0), ("gr_skewed_128", 1), ("g2_uniform_128", 2), ("g1_dup_neg_96", 0), ("g2_identity_bases_24", 2),
             ("gr_uniform_100", 1), ("g1_uniform_0", 0), ("g2_dup_neg_48", 2), ("g1_identity_bases_40", 0)]
    jobs, expect, keep = [], [], []
    for i, (name, curve) in enumerate(specs):
        z = golden_g2 if curve == 2 else golden
        pts, sc, inf, aff = (z[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
        bases = k.Bases(curve, pts, inf)
        if i == 3:
            bases.precompute(6)
        keep.append(bases)
        jobs.append((bases, sc))
        expect.append((curve, aff))
    # an offset window and a canonical-format job on vectors already in the batch
    pts, sc = golden["g1_uniform_1024_pts"], golden["g1_uniform_1024_sc"]
    jobs.append((keep[0], sc[:1000], 24))
    expect.append((0, A.to_affine(0, A.msm(0, pts[24:], sc[:1000]))))
    canon = np.stack([A.field_op(A.FIELD_FR, "mont_reduce", s) for s in golden_g2["g2_uniform_128_sc"]])
    jobs.append((keep[2], canon, 0, k.SCALARS_CANONICAL))
    expect.append((2, golden_g2["g2_uniform_128_aff"]))
    for _ in range(2):
        got = k.msm_batch(jobs)
        assert len(got) == len(jobs)
        for g, (curve, aff) in zip(got, expect):
            assert same_affine(k.to_affine(curve, g), aff)
    for b in keep:
        b.free()
