"""Oracle for row N4 (nova folding vector work): restated SparseMatrix::prod / compute_cross_term / fold against big-int
arithmetic and against the property the reference's own tests check — folding preserves relaxed-R1CS satisfiability
(nova/src/prover.rs tests `folding_scheme_prover_test`, relaxed_r1cs.rs:80-114 `is_sat_relaxed`)."""
import numpy as np
import pytest
from nova_util import csr_from_rows, example_r1cs, example_z, ints, mont, relaxed_sat

from oracle import oracle as A
from oracle import pyref as B

FIELDS = [(A.FIELD_FQ, B.FQ), (A.FIELD_FR, B.FR)]


@pytest.mark.parametrize("fid,p", FIELDS)
def test_sparse_prod_matches_bigint(fid, p):
    rng = np.random.default_rng(5)
    m, n_z = 57, 23
    rows = []
    for i in range(m):
        k = [0, 1, 2, 5, n_z][i % 5]
        cols = rng.choice(n_z, size=k, replace=False)
        rows.append({int(c): int(rng.integers(1, 1 << 62)) ** 4 % p if i % 3 else 1 for c in cols})
    mat = csr_from_rows(rows, p)
    z_int = [int(rng.integers(0, 1 << 62)) ** 4 % p for _ in range(n_z)]
    got = ints(A.sparse_prod(fid, m, mat, mont(z_int, p)), p)
    assert got == [sum(co * z_int[c] for c, co in r.items()) % p for r in rows]


@pytest.mark.parametrize("fid,p", FIELDS)
def test_example_r1cs_is_satisfied_and_folding_keeps_it_satisfied(fid, p):
    m, n_z, mats = example_r1cs(p)
    prod = lambda k, z: A.sparse_prod(fid, m, mats[k], z)
    for x in (0, 3, 4, 12345):
        assert relaxed_sat(prod, m, mats, example_z(x, p), [0] * m, p)                    # is_sat (relaxed_r1cs.rs:117-146) with u = 1, E = 0
    assert not relaxed_sat(prod, m, mats, example_z(3, p)[:-1] + [7], [0] * m, p)
    # running instance (u, x, w, E), folded twice with fresh instances (prover.rs:24-50 without the transcript)
    z1, e1 = example_z(3, p), [0] * m
    for x2, r in ((4, 0x1234567890ABCDEF1234567890ABCDEF % p), (9, p - 5)):
        z2 = example_z(x2, p)
        t = A.cross_term(fid, m, *mats, mont(z1, p), mont(z2, p))
        # T by the definition, with big integers
        az1, bz1, cz1 = (ints(A.sparse_prod(fid, m, mats[k], mont(z1, p)), p) for k in range(3))
        az2, bz2, cz2 = (ints(A.sparse_prod(fid, m, mats[k], mont(z2, p)), p) for k in range(3))
        assert ints(t, p) == [(az2[i] * bz1[i] + az1[i] * bz2[i] - z1[0] * cz2[i] - z2[0] * cz1[i]) % p for i in range(m)]
        rm = mont([r], p)[0]
        z1 = ints(A.vec_fold(fid, mont(z1, p), mont(z2, p), rm), p)                        # u1 + r, x1 + r x2, w1 + r w2 (instance.rs:88-93, witness.rs:68)
        e1 = ints(A.vec_fold(fid, mont(e1, p), t, rm), p)                                  # e1 + r t (witness.rs:67)
        assert z1[0] == (1 + r) % p or x2 == 9
        assert relaxed_sat(prod, m, mats, z1, e1, p)
    assert any(e1)                                                                          # the error term is non-trivial after folding
