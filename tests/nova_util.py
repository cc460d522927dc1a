"""Helpers for the Nova folding tests: CSR builders and the relaxed-R1CS satisfiability check (nova/src/relaxed_r1cs.rs:80-114)."""
import numpy as np

from oracle import pyref as B


def mont(vals, p):
    return np.array([B.int_to_limbs(B.to_mont(v % p, p)) for v in vals], dtype=np.uint64).reshape(-1, 4)


def ints(arr, p):
    return [B.from_mont(B.limbs_to_int(x), p) for x in np.asarray(arr, dtype=np.uint64).reshape(-1, 4)]


def csr_from_rows(rows, p):
    """rows: list of {column: coeff} -> (row_ptr, cols, coeffs Montgomery)."""
    row_ptr, cols, co = [0], [], []
    for r in rows:
        for c, v in r.items():
            if v % p:
                cols.append(c)
                co.append(v)
        row_ptr.append(len(cols))
    return np.array(row_ptr, dtype=np.uint32), np.array(cols, dtype=np.uint32), mont(co, p)


def example_r1cs(p):
    """zkstd/src/r1cs/test.rs:54-95: x^3 + x + 5 = out as dense 4 x 6 matrices over z = (1, input, out, x^2, x^3, x^3 + x).
    dense_to_sparse (:14-36) maps column j <= l to Wire::Instance(j) and the rest to Wire::Witness(j - l - 1); nova resolves
    them back to z[j] (matrix.rs:41-44 with l = x.len() + 1, relaxed_r1cs.rs:96-97), so the flat column is j."""
    a = [[0, 1, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 1, 0, 0, 1, 0], [5, 0, 0, 0, 0, 1]]
    b = [[0, 1, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]]
    c = [[0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1], [0, 0, 1, 0, 0, 0]]
    to_rows = lambda d: [{j: v for j, v in enumerate(r) if v} for r in d]
    return 4, 6, tuple(csr_from_rows(to_rows(d), p) for d in (a, b, c))


def example_z(x, p):
    """test.rs:38-52 example_z_witness, as the whole z = (1, x, w)."""
    return [1, x % p, (x ** 3 + x + 5) % p, x * x % p, x ** 3 % p, (x ** 3 + x) % p]


def chain_r1cs(steps, x0, p):
    """The example's function iterated (the map nova/src/test.rs:16-29 folds): x_{i+1} = x_i^3 + x_i + 5, three constraints per
    step plus one tying the public output; z = (1, x0, out, w...).  The shape does not depend on x0.  -> m, n_z, (A, B, C), z"""
    z = [1, x0 % p, None]
    a, b, c = [], [], []
    cur = 1
    for _ in range(steps):
        x = z[cur]
        sq = len(z); z.append(x * x % p); a.append({cur: 1}); b.append({cur: 1}); c.append({sq: 1})
        cu = len(z); z.append(z[sq] * x % p); a.append({sq: 1}); b.append({cur: 1}); c.append({cu: 1})
        nx = len(z); z.append((z[cu] + x + 5) % p); a.append({cu: 1, cur: 1, 0: 5}); b.append({0: 1}); c.append({nx: 1})
        cur = nx
    z[2] = z[cur]
    a.append({cur: 1}); b.append({0: 1}); c.append({2: 1})
    return len(a), len(z), tuple(csr_from_rows(r, p) for r in (a, b, c)), z


def relaxed_sat(prod, m, mats, z_int, e_int, p):
    """(A z) o (B z) == u (C z) + E with u = z[0]; `prod(which, z_mont)` is the matrix-vector product under test."""
    z = mont(z_int, p)
    az, bz, cz = (ints(prod(k, z), p) for k in range(3))
    u = z_int[0]
    return all((az[i] * bz[i] - u * cz[i] - e_int[i]) % p == 0 for i in range(m))
