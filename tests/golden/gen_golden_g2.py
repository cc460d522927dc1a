"""Generates tests/golden/g2_vectors.npz — run from the repo root:  python tests/golden/gen_golden_g2.py

BN254 G2 cases for the next-row G2 MSM (groth16/src/prover.rs:64-65).  Like msm_vectors.npz the values are
minted from Oracle A (oracle/zkstd_oracle.cpp: the reference's generic point formulas over Fq2, bn254/src/fqn.rs)
and every case is cross-checked against textbook big-int affine arithmetic over Fq2 (oracle/groth16_ref.py g2_*)
before it is written.  Points are (n, 16) uint64 = x.c0 x.c1 y.c0 y.c1, Montgomery limbs.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import groth16_ref as G  # noqa: E402
from oracle import oracle as A  # noqa: E402
from oracle import pyref as B  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "g2_vectors.npz")
Q, R = B.FQ, B.FR
CURVE = A.BN254_G2


def dec2(l):
    return (B.from_mont(B.limbs_to_int(l[:4]), Q), B.from_mont(B.limbs_to_int(l[4:8]), Q))


def bigint_check(pts, inf, scalars, expect_aff):
    exp = None
    for p, f, s in zip(pts, inf, scalars):
        if f:
            continue
        exp = G.g2_add(exp, G.g2_mul((dec2(p[:8]), dec2(p[8:])), B.from_mont(B.limbs_to_int(s), R)))
    if exp is None:
        assert int(expect_aff[16]) == 1
    else:
        assert int(expect_aff[16]) == 0 and (dec2(expect_aff[:8]), dec2(expect_aff[8:16])) == exp, "Oracle A != big-int G2"


def neg_y(p):
    q = p.copy()
    for h in (8, 12):
        y = B.limbs_to_int(q[h:h + 4])
        q[h:h + 4] = B.int_to_limbs((Q - y) % Q)
    return q


def main():
    cases = {}
    pool = A.random_points(CURVE, 128, seed=A.DEFAULT_SEED)
    sc_pool = A.random_field(A.FIELD_FR, 128, seed=bytes(reversed(A.DEFAULT_SEED)))
    rm1 = np.array(B.int_to_limbs(B.to_mont(R - 1, R)), dtype=np.uint64)
    one = np.array(B.int_to_limbs(B.to_mont(1, R)), dtype=np.uint64)

    def add(name, pts, scalars, inf=None):
        pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 16)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        inf_a = np.zeros(pts.shape[0], dtype=np.uint8) if inf is None else np.asarray(inf, dtype=np.uint8)
        aff = A.to_affine(CURVE, A.msm(CURVE, pts, scalars, inf=inf_a))
        n = min(pts.shape[0], scalars.shape[0])
        bigint_check(pts[:n], inf_a[:n], scalars[:n], aff)
        key = f"g2_{name}"
        cases[key + "_pts"], cases[key + "_sc"], cases[key + "_inf"], cases[key + "_aff"] = pts, scalars, inf_a, aff
        print(key, "n_bases", pts.shape[0], "n_scalars", scalars.shape[0], "inf_result", int(aff[16]))

    for n in (0, 1, 3, 33, 128):
        add(f"uniform_{n}", pool[:n], sc_pool[:n])
    add("more_bases", pool[:48], sc_pool[:20])
    add("more_scalars", pool[:20], sc_pool[:48])
    add("zero_scalars", pool[:32], np.zeros((32, 4), dtype=np.uint64))
    add("rm1_scalars", pool[:32], np.tile(rm1, (32, 1)))
    sk = sc_pool[:64].copy()
    sk[0::2] = 0
    sk[1::4] = one
    add("skewed_64", pool[:64], sk)
    dup, scd = pool[:48].copy(), sc_pool[:48].copy()
    for i in range(0, 48, 3):
        dup[i + 1], scd[i + 1] = dup[i], scd[i]
        dup[i + 2], scd[i + 2] = neg_y(dup[i]), scd[i]
    add("dup_neg_48", dup, scd)
    canc = np.concatenate([pool[:8], np.stack([neg_y(p) for p in pool[:8]])])
    add("cancel_16", canc, np.concatenate([sc_pool[:8], sc_pool[:8]]))
    inf = np.zeros(24, dtype=np.uint8)
    inf[[0, 7, 8, 23]] = 1
    idp = pool[:24].copy()
    for i in np.nonzero(inf)[0]:
        idp[i] = 0
        idp[i, 8:12] = np.array(B.int_to_limbs(B.to_mont(1, Q)), dtype=np.uint64)  # (0, 1, inf), group.rs:22-26
    add("identity_bases_24", idp, sc_pool[:24], inf)
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
