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


@pytest.mark.parametrize("curve", [A.GRUMPKIN, A.BN254_G1])
def test_resident_folding_step(k, curve):
    """One NIFS step with z1, z2, E, T resident on the GPU (witness.rs:20-21,56-71; prover.rs:33-35; ivc.rs:160-205): the witness commitment
    from z[1 + l ..], the cross term and its commitment, and both folds never cross PCIe — every value equals the host-buffer path / the
    oracle, and the folded (z, E) still satisfies the relaxed relation."""
    from kogarashi_b200 import nova
    fid = A.SCALAR_FIELD[curve]
    p = B.FQ if fid == A.FIELD_FQ else B.FR
    m, n_z, mats, z1_int = chain_r1cs(700, 3, p)
    z2_int = chain_r1cs(700, 5, p)[3]
    z1, z2 = mont(z1_int, p), mont(z2_int, p)
    l = 3                                                              # z = (u, x0, out | w)
    shape = nova.R1csShape(fid, m, n_z, *mats)
    ck = k.PedersenCommitment(curve, A.random_points(curve, max(m, n_z), seed=bytes(range(5, 21))))
    d_z1, d_z2 = nova.DeviceVec(fid, z1), nova.DeviceVec(fid, z2)
    d_e, d_t = nova.DeviceVec(fid, n=m), nova.DeviceVec(fid, n=m)
    assert len(d_z1) == n_z and (d_z1.download() == z1).all() and (d_e.download() == 0).all()
    # commit(W) of the incoming witness straight from the resident z (R1csWitness::commit, relaxed_r1cs.rs:36)
    assert same_affine(d_z2.commit(ck, off=l), ck.commit(z2[l:]))
    assert same_affine(d_z2.commit(ck, off=l, n=100), ck.commit(z2[l:l + 100]))
    # T and commit(T)
    commit_t = nova.cross_term_device(shape, d_z1, d_z2, t=d_t, ck=ck)
    t_ref = A.cross_term(fid, m, *mats, z1, z2)
    assert (d_t.download() == t_ref).all()
    assert same_affine(commit_t, ck.commit(t_ref))
    assert nova.cross_term_device(shape, d_z1, d_z2, t=d_t) is None     # without a key: T only
    # folds in place: z <- z1 + r z2, E <- E + r T
    r = 0x1F2E3D4C5B6A79880123456789ABCDEF % p
    rm = mont([r], p)[0]
    d_z1.fold(d_z2, rm)
    d_e.fold(d_t, rm)
    assert (d_z1.download() == A.vec_fold(fid, z1, z2, rm)).all()
    assert (d_e.download() == A.vec_fold(fid, np.zeros_like(t_ref), t_ref, rm)).all()
    assert relaxed_sat(shape.prod, m, mats, ints(d_z1.download(), p), ints(d_e.download(), p), p)
    # a second step on the folded state, with the fresh instance written into the resident z2
    z3 = mont(chain_r1cs(700, 7, p)[3], p)
    d_z2.write(0, z3)
    zf, ef = d_z1.download(), d_e.download()
    commit_t2 = nova.cross_term_device(shape, d_z1, d_z2, t=d_t, ck=ck)
    t2_ref = A.cross_term(fid, m, *mats, zf, z3)
    assert same_affine(commit_t2, ck.commit(t2_ref))
    d_z1.fold(d_z2, rm)
    d_e.fold(d_t, rm)
    assert relaxed_sat(shape.prod, m, mats, ints(d_z1.download(), p), ints(d_e.download(), p), p)
    # errors: wrong field, short vector, range
    other = nova.DeviceVec(1 - fid, n=n_z)
    with pytest.raises(k.KgrError):
        nova.cross_term_device(shape, d_z1, other)
    with pytest.raises(k.KgrError):
        d_z1.fold(other, rm)
    with pytest.raises(k.KgrError):
        nova.cross_term_device(shape, d_z1, nova.DeviceVec(fid, n=n_z - 1))
    with pytest.raises(k.KgrError):
        d_z1.download(n_z - 1, 2)
    for v in (d_z1, d_z2, d_e, d_t, other):
        v.free()
    shape.free()
