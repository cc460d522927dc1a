"""Oracle A on BN254 G2 (Fq2 coordinates, bn254/src/fqn.rs + g2.rs through the reference's generic point formulas)
against textbook big-int arithmetic (oracle/groth16_ref.py) and the committed G2 goldens."""
import numpy as np
import pytest
from conftest import golden_g2_case_names, same_affine

from oracle import groth16_ref as G
from oracle import oracle as A
from oracle import pyref as B

Q, R = B.FQ, B.FR
C2 = A.BN254_G2


def dec2(l):
    return (B.from_mont(B.limbs_to_int(l[:4]), Q), B.from_mont(B.limbs_to_int(l[4:8]), Q))


def dec_pt(a):
    return None if int(a[16]) else (dec2(a[:8]), dec2(a[8:16]))


def test_generator_and_group_law_match_bigint():
    g = A.generator(C2)
    assert (dec2(g[:8]), dec2(g[8:])) == G.G2_GEN                     # bn254/src/params.rs:15-42
    assert int(A.point_op(C2, 6, np.concatenate([g, np.zeros(1, np.uint64)]), out_len=1)[0]) == 1
    one = np.array(B.int_to_limbs(B.to_mont(1, Q)), dtype=np.uint64)
    proj = np.concatenate([g, one, np.zeros(4, np.uint64)])
    assert dec_pt(A.to_affine(C2, A.point_op(C2, 1, proj))) == G.g2_add(G.G2_GEN, G.G2_GEN)
    for k in (1, 2, 3, 5, R - 1, 0x1234567890ABCDEF1234567890ABCDEF):
        s = np.array(B.int_to_limbs(B.to_mont(k % R, R)), dtype=np.uint64)
        assert dec_pt(A.to_affine(C2, A.scalar_point(C2, proj, s))) == G.g2_mul(G.G2_GEN, k)
    # k = r gives the identity: the generator has order r
    assert int(A.to_affine(C2, A.scalar_point(C2, proj, np.zeros(4, np.uint64)))[16]) == 1


def test_random_points_are_k_times_generator():
    xy, ks = A.random_points(C2, 6, return_scalars=True)
    for p, k in zip(xy, ks):
        assert (dec2(p[:8]), dec2(p[8:])) == G.g2_mul(G.G2_GEN, B.from_mont(B.limbs_to_int(k), R))


@pytest.mark.parametrize("name", golden_g2_case_names())
def test_oracle_reproduces_g2_goldens(golden_g2, name):
    pts, sc, inf, aff = (golden_g2[name + s] for s in ("_pts", "_sc", "_inf", "_aff"))
    for threads in (1, 3):
        assert same_affine(A.to_affine(C2, A.msm(C2, pts, sc, inf=inf, threads=threads)), aff)


def test_msm_equals_scalar_times_generator_checksum():
    """Bases k_i * G: msm must equal (sum k_i s_i) * G (the size-independent property the GPU tests use at 2^18)."""
    n = 300
    xy, ks = A.random_points(C2, n, seed=bytes(range(16)), return_scalars=True)
    sc = A.random_field(A.FIELD_FR, n, seed=bytes(range(1, 17)))
    tot = sum(B.from_mont(B.limbs_to_int(a), R) * B.from_mont(B.limbs_to_int(b), R) for a, b in zip(ks, sc)) % R
    assert dec_pt(A.to_affine(C2, A.msm(C2, xy, sc))) == G.g2_mul(G.G2_GEN, tot)
