"""GPU tests of the in-process multi-GPU path (kgr_init with several devices): contiguous shards, one partial
result per GPU, host-side combination.  Skipped on a single-GPU box."""
import numpy as np
import pytest
from conftest import same_affine

from oracle import oracle as A

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("curve", [A.BN254_G1, A.GRUMPKIN])
def test_sharded_msm_matches_oracle(curve):
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    import kogarashi_b200 as k
    k.init(list(range(min(ng, 8))))
    try:
        n = 5001  # ragged shards
        pool = A.random_points(curve, 512, seed=bytes(range(3, 19)))
        pts = np.tile(pool, (10, 1))[:n]
        sc = A.random_field(A.SCALAR_FIELD[curve], n, seed=bytes(range(4, 20)))
        exp = A.to_affine(curve, A.msm(curve, pts, sc))
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(pts, sc, curve=curve)), exp)          # oneshot, sharded
        bases = k.Bases(curve, pts)
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc)), exp)                      # registered, sharded
        off = 1234                                                                                        # window crossing shard borders
        exp2 = A.to_affine(curve, A.msm(curve, pts[off:], sc[: n - off]))
        assert same_affine(k.to_affine(curve, k.msm_curve_addition(bases, sc[: n - off], base_off=off)), exp2)
        tiny = k.msm_curve_addition(pts[:3], sc[:3], curve=curve)                                         # fewer pairs than GPUs
        assert same_affine(k.to_affine(curve, tiny), A.to_affine(curve, A.msm(curve, pts[:3], sc[:3])))
        bases.free()
    finally:
        k.init()


def test_g2_and_batch_on_several_devices():
    """G2 through the sharded path, and kgr_msm_batch with several devices selected (jobs run in sequence, each spread over all GPUs)."""
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    import kogarashi_b200 as k
    k.init(list(range(min(ng, 8))))
    try:
        n = 3001
        p2 = np.tile(A.random_points(A.BN254_G2, 256, seed=bytes(range(3, 19))), (12, 1))[:n]
        p1 = np.tile(A.random_points(A.BN254_G1, 256, seed=bytes(range(5, 21))), (12, 1))[:n]
        sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(4, 20)))
        e2 = A.to_affine(A.BN254_G2, A.msm(A.BN254_G2, p2, sc))
        e1 = A.to_affine(A.BN254_G1, A.msm(A.BN254_G1, p1, sc))
        assert same_affine(k.to_affine(A.BN254_G2, k.msm_curve_addition(p2, sc, curve=A.BN254_G2)), e2)
        b2, b1 = k.Bases(A.BN254_G2, p2), k.Bases(A.BN254_G1, p1)
        got = k.msm_batch([(b2, sc), (b1, sc), (b1, sc[:100], 7)])
        assert same_affine(k.to_affine(A.BN254_G2, got[0]), e2) and same_affine(k.to_affine(A.BN254_G1, got[1]), e1)
        assert same_affine(k.to_affine(A.BN254_G1, got[2]), A.to_affine(A.BN254_G1, A.msm(A.BN254_G1, p1[7:], sc[:100])))
        b1.free()
        b2.free()
    finally:
        k.init()


def test_pieces_on_every_device():
    """A sharded call large enough for pipelined pieces on each device (forced piece counts and the automatic rule), registered and oneshot."""
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    import kogarashi_b200 as k
    g = min(ng, 8)
    k.init([0])
    n = (1 << 19) + 333
    one = k.Bases.generate(A.BN254_G1, n, seed=17)
    pts = one.download()
    sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(11, 27)))
    ref = k.to_affine(A.BN254_G1, k.msm_curve_addition(one, sc))
    one.free()
    k.init(list(range(g)))
    try:
        bases = k.Bases(A.BN254_G1, pts)
        for pieces in (0, 1, 2, 3):
            k.set_param("oneshot_split", pieces)
            assert same_affine(k.to_affine(A.BN254_G1, k.msm_curve_addition(bases, sc)), ref), pieces
            assert same_affine(k.to_affine(A.BN254_G1, k.msm_curve_addition(pts, sc, curve=A.BN254_G1)), ref), pieces
            assert same_affine(k.to_affine(A.BN254_G1, k.msm_curve_addition(bases, sc[: n - 1000], base_off=1000)),
                               k.to_affine(A.BN254_G1, k.msm_curve_addition(pts[1000:], sc[: n - 1000], curve=A.BN254_G1))), pieces
        bases.free()
    finally:
        k.set_param("oneshot_split", 0)
        k.init()


def test_groth16_msms_on_several_devices():
    """kgr_groth16_msms with several devices selected takes the plain sequence (H, h query, batch): same results as the fused single-device call."""
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    import kogarashi_b200 as k
    from kogarashi_b200.fft import Fft
    kk, m = 10, 1000
    ev = [A.random_field(A.FIELD_FR, m, seed=bytes(range(i, i + 16))) for i in (1, 2, 3)]
    hp = A.random_points(A.BN254_G1, (1 << kk) - 1, seed=bytes(range(5, 21)))
    zp = A.random_points(A.BN254_G1, 300, seed=bytes(range(6, 22)))
    z = A.random_field(A.FIELD_FR, 300, seed=bytes(range(7, 23)))
    results = []
    for devs in ([0], list(range(min(ng, 8)))):
        k.init(devs)
        h, zb = k.Bases(A.BN254_G1, hp), k.Bases(A.BN254_G1, zp)
        q_pt, q, res = k.groth16_msms(kk, *ev, h, [(zb, z)])
        q_ref = Fft(kk).h_coefficients(*ev)
        assert q.shape == q_ref.shape and (q == q_ref).all()
        assert same_affine(k.to_affine(A.BN254_G1, q_pt), k.to_affine(A.BN254_G1, k.msm_curve_addition(h, q_ref)))
        results.append((k.to_affine(A.BN254_G1, q_pt), k.to_affine(A.BN254_G1, res[0])))
        h.free()
        zb.free()
    k.init()
    assert same_affine(results[0][0], results[1][0]) and same_affine(results[0][1], results[1][1])
    assert same_affine(results[0][1], A.to_affine(A.BN254_G1, A.msm(A.BN254_G1, zp, z)))
