"""GPU parity tests for row N4: sparse R1CS products, Nova's cross term (+ its commitment from device-resident scalars) and
the witness fold, through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest
from conftest import same_affine
from nova_util import chain_r1cs, csr_from_rows, example_r1cs, example_z, ints, mont, relaxed_sat

from oracle import oracle as A
from oracle import pyref as B

pytestmark = pytest.mark.gpu
FIELDS = [(A.FIELD_FQ, B.FQ), (A.FIELD_FR, B.FR)]


@pytest.fixture(scope="module")
def k():
    import kogarashi_b200 as kk
    kk.init()
    return kk


def _random_shape(rng, m, n_z, p):
    mats = []
    for _ in range(3):
        rows = []
        for i in range(m):
            kind = rng.integers(0, 10)
            nnz = 0 if kind == 0 else (min(n_z, 300) if (kind == 1 and i % 97 == 0) else int(rng.integers(1, 5)))
            cols = rng.choice(n_z, size=nnz, replace=False)
            rows.append({int(c): (1 if rng.integers(0, 2) else int(rng.integers(1, 1 << 62)) ** 4 % p) for c in cols})
        mats.append(csr_from_rows(rows, p))
    return tuple(mats)


@pytest.mark.parametrize("fid,p", FIELDS)
def test_prod_cross_term_fold_bit_exact(k, fid, p):
    from kogarashi_b200 import nova
    rng = np.random.default_rng(11 + fid)
    m, n_z = 1500, 700
    mats = _random_shape(rng, m, n_z, p)
    z1 = A.random_field(fid, n_z, seed=bytes(range(16)))
    z2 = A.random_field(fid, n_z, seed=bytes(range(1, 17)))
    shape = nova.R1csShape(fid, m, n_z, *mats)
    for which in range(3):
        assert (shape.prod(which, z1) == A.sparse_prod(fid, m, mats[which], z1)).all()
    t = shape.cross_term(z1, z2)
    assert (t == A.cross_term(fid, m, *mats, z1, z2)).all()
    r = A.random_field(fid, 1, seed=bytes(range(2, 18)))[0]
    assert (nova.vec_fold(fid, z1, z2, r) == A.vec_fold(fid, z1, z2, r)).all()
    assert nova.vec_fold(fid, z1[:0], z2[:0], r).shape == (0, 4)
    shape.free()


@pytest.mark.parametrize("fid,p", FIELDS)
def test_example_r1cs_fold_keeps_relaxed_satisfiability(k, fid, p):
    """zkstd/src/r1cs/test.rs example through the device kernels: T, folded z and E satisfy (A z) o (B z) = u (C z) + E."""
    from kogarashi_b200 import nova
    m, n_z, mats = example_r1cs(p)
    shape = nova.R1csShape(fid, m, n_z, *mats)
    z1, e1 = example_z(3, p), [0] * m
    assert relaxed_sat(shape.prod, m, mats, z1, e1, p)
    for x2, r in ((4, 0xDEADBEEFCAFEF00D1234 % p), (9, p - 2)):
        z2 = example_z(x2, p)
        t = shape.cross_term(mont(z1, p), mont(z2, p))
        assert (t == A.cross_term(fid, m, *mats, mont(z1, p), mont(z2, p))).all()
        rm = mont([r], p)[0]
        z1 = ints(nova.vec_fold(fid, mont(z1, p), mont(z2, p), rm), p)
        e1 = ints(nova.vec_fold(fid, mont(e1, p), t, rm), p)
        assert relaxed_sat(shape.prod, m, mats, z1, e1, p)
    shape.free()


@pytest.mark.parametrize("curve", [A.GRUMPKIN, A.BN254_G1])
def test_cross_term_commitment_from_device_scalars(k, curve):
    """prover.rs:33-35: t = compute_cross_term(..); commit_t = ck.commit(&t) — one call, T never leaves the device for the MSM.
    Chained x^3 + x + 5 circuit (4096 constraints) with two different inputs."""
    from kogarashi_b200 import nova
    fid = A.SCALAR_FIELD[curve]
    p = B.FQ if fid == A.FIELD_FQ else B.FR
    m, n_z, mats, z1_int = chain_r1cs(1365, 3, p)
    z2_int = chain_r1cs(1365, 4, p)[3]
    z1, z2 = mont(z1_int, p), mont(z2_int, p)
    shape = nova.R1csShape(fid, m, n_z, *mats)
    ck = k.PedersenCommitment(curve, A.random_points(curve, 4096, seed=bytes(range(7, 23))))
    t, commit = shape.cross_term(z1, z2, ck=ck)
    t_ref = A.cross_term(fid, m, *mats, z1, z2)
    assert (t == t_ref).all()
    assert same_affine(commit, ck.commit(t_ref))
    tm = shape.last_timing()
    assert tm["cross_term"] > 0 and tm["commit"] > 0
    # fold and check the relaxed relation with the device products
    r = 0x0123456789ABCDEF0123456789ABCDEF % p
    rm = mont([r], p)[0]
    zf = ints(nova.vec_fold(fid, z1, z2, rm), p)
    ef = ints(nova.vec_fold(fid, np.zeros_like(t), t, rm), p)
    assert relaxed_sat(shape.prod, m, mats, zf, ef, p)
    # a key on the wrong curve is rejected
    other = A.GRUMPKIN if curve == A.BN254_G1 else A.BN254_G1
    bad = k.PedersenCommitment(other, A.random_points(other, 8))
    with pytest.raises(k.KgrError):
        shape.cross_term(z1, z2, ck=bad)
    shape.free()


def test_register_rejects_malformed_csr(k):
    from kogarashi_b200 import nova
    p = B.FR
    good = csr_from_rows([{0: 1}, {1: 2}], p)
    bad_col = (good[0], np.array([0, 5], dtype=np.uint32), good[2])
    with pytest.raises(k.KgrError):
        nova.R1csShape(A.FIELD_FR, 2, 2, good, good, bad_col)
    bad_ptr = (np.array([0, 2, 1], dtype=np.uint32), good[1], good[2])
    with pytest.raises(AssertionError):
        nova.R1csShape(A.FIELD_FR, 2, 2, bad_ptr, good, good)
