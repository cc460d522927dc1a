"""e2e of a streamed host-buffer call against (pieces, affine levels): python tools/probe_e2e_levels.py [logn]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
bases = k.Bases.generate(0, n, seed=3)
pts = torch.from_numpy(bases.download().view(np.int64)).pin_memory()
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
scp = torch.from_numpy(sc.view(np.int64)).pin_memory()
ref = k.to_affine(0, k.msm_curve_addition(bases, sc))
for levels in (-1, 0, 1, 2, 3):
    k.set_param("affine_levels", levels)
    row = []
    for pieces in (1, 2, 3, 4):
        k.set_param("oneshot_split", pieces)
        for _ in range(3):
            out = k.msm_oneshot_ptr(0, pts.data_ptr(), n, scp.data_ptr(), n)
        t0 = time.perf_counter()
        for _ in range(10):
            out = k.msm_oneshot_ptr(0, pts.data_ptr(), n, scp.data_ptr(), n)
        dt = (time.perf_counter() - t0) / 10
        assert (k.to_affine(0, out) == ref).all()
        for _ in range(3):
            out = k.msm_host_ptr(bases, scp.data_ptr(), n)
        t0 = time.perf_counter()
        for _ in range(10):
            out = k.msm_host_ptr(bases, scp.data_ptr(), n)
        dr = (time.perf_counter() - t0) / 10
        row.append(f"p{pieces}: {dt*1e3:.2f}/{dr*1e3:.2f}")
    print(f"2^{logn} affine_levels={levels}: oneshot/registered ms  " + "  ".join(row), flush=True)
