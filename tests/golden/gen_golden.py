"""Generates tests/golden/msm_vectors.npz — run from the repo root:  python tests/golden/gen_golden.py

The reference ships no fixed MSM vectors and cannot be executed here (no Rust
toolchain), so the golden values are minted from Oracle A (oracle/zkstd_oracle.cpp, the restated
reference algorithm) and every case is cross-checked against Oracle B (oracle/pyref.py, textbook
big-int affine arithmetic) before it is written.  Inputs come from the reference's own sampler
(from_u512 of 8 next_u64, zkstd .../represent.rs:80-103) on the restated xorshift128 stream seeded
as in pallet/nova/src/tests.rs:69-74.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle as A  # noqa: E402
from oracle import pyref as B  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "msm_vectors.npz")


def pyref_check(curve, pts, inf, scalars, expect_aff):
    cm = B.CURVES[curve]
    P = []
    for p, f in zip(pts, inf):
        P.append(None if f else (B.from_mont(B.limbs_to_int(p[:4]), cm.p), B.from_mont(B.limbs_to_int(p[4:]), cm.p)))
    S = [B.from_mont(B.limbs_to_int(s), cm.r) for s in scalars]
    exp = cm.msm(P, S)
    if exp is None:
        assert int(expect_aff[8]) == 1
    else:
        got = (B.from_mont(B.limbs_to_int(expect_aff[:4]), cm.p), B.from_mont(B.limbs_to_int(expect_aff[4:8]), cm.p))
        assert int(expect_aff[8]) == 0 and got == exp, "Oracle A != Oracle B"


def main():
    cases = {}
    pool_n = 1024
    for curve, cname in ((A.BN254_G1, "g1"), (A.GRUMPKIN, "gr")):
        sf = A.SCALAR_FIELD[curve]
        cm = B.CURVES[curve]
        pool = A.random_points(curve, pool_n, seed=A.DEFAULT_SEED)
        sc_pool = A.random_field(sf, pool_n, seed=bytes(reversed(A.DEFAULT_SEED)))
        rm1 = np.array(B.int_to_limbs(B.to_mont(cm.r - 1, cm.r)), dtype=np.uint64)
        one = np.array(B.int_to_limbs(B.to_mont(1, cm.r)), dtype=np.uint64)

        def add(name, pts, scalars, inf=None):
            pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 8)
            scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
            inf_a = np.zeros(pts.shape[0], dtype=np.uint8) if inf is None else np.asarray(inf, dtype=np.uint8)
            proj = A.msm(curve, pts, scalars, inf=inf_a)
            aff = A.to_affine(curve, proj)
            n = min(pts.shape[0], scalars.shape[0])
            if n <= 160:
                pyref_check(curve, pts[:n], inf_a[:n], scalars[:n], aff)
            key = f"{cname}_{name}"
            cases[key + "_pts"] = pts
            cases[key + "_sc"] = scalars
            cases[key + "_inf"] = inf_a
            cases[key + "_aff"] = aff
            print(key, "n_bases", pts.shape[0], "n_scalars", scalars.shape[0], "inf_result", int(aff[8]))

        for n in (0, 1, 2, 3, 4, 31, 32, 33, 100):
            add(f"uniform_{n}", pool[:n], sc_pool[:n])
        add("uniform_1024", pool, sc_pool)
        # len(bases) != len(coeffs): the prover relies on zip semantics (groth16/src/prover.rs:58)
        add("more_bases", pool[:48], sc_pool[:20])
        add("more_scalars", pool[:20], sc_pool[:48])
        n = 64
        add("zero_scalars", pool[:n], np.zeros((n, 4), dtype=np.uint64))
        add("rm1_scalars", pool[:n], np.tile(rm1, (n, 1)))
        # skewed: 50% zeros, 25% ones, rest uniform (Nova witness shape)
        sk = sc_pool[:128].copy()
        sk[0::2] = 0
        sk[1::4] = one
        add("skewed_128", pool[:128], sk)
        # duplicated points and P / -P pairs with equal scalars (bucket doubling / cancellation paths)
        dup = pool[:96].copy()
        scd = sc_pool[:96].copy()
        for i in range(0, 96, 3):
            dup[i + 1] = dup[i]
            scd[i + 1] = scd[i]
            neg = dup[i].copy()
            y = B.limbs_to_int(neg[4:])
            neg[4:] = B.int_to_limbs((cm.p - y) % cm.p)
            dup[i + 2] = neg
            scd[i + 2] = scd[i]
        add("dup_neg_96", dup, scd)
        # everything cancels -> identity result
        canc = np.concatenate([pool[:16], pool[:16].copy()])
        for i in range(16):
            y = B.limbs_to_int(canc[16 + i][4:])
            canc[16 + i][4:] = B.int_to_limbs((cm.p - y) % cm.p)
        add("cancel_32", canc, np.concatenate([sc_pool[:16], sc_pool[:16]]))
        # identity bases (Groth16 CRS entries left at ADDITIVE_IDENTITY, groth16/src/zksnark.rs:62-66)
        inf = np.zeros(40, dtype=np.uint8)
        inf[[0, 7, 8, 39]] = 1
        idp = pool[:40].copy()
        for i in np.nonzero(inf)[0]:
            idp[i, :4] = 0
            idp[i, 4:] = np.array(B.int_to_limbs(B.to_mont(1, cm.p)), dtype=np.uint64)  # (0, 1, inf) as in group.rs:22-26
        add("identity_bases_40", idp, sc_pool[:40], inf)
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
