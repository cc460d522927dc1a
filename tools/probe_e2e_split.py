"""e2e (kgr_msm_oneshot from pinned host memory) against the number of pipelined pieces: python tools/probe_e2e_split.py [logn] [curve]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << logn
bases = k.Bases.generate(curve, n, seed=3)
pts = torch.from_numpy(bases.download().view(np.int64)).pin_memory()
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
scp = torch.from_numpy(sc.view(np.int64)).pin_memory()
ref = k.to_affine(curve, k.msm_curve_addition(bases, sc))
for pieces in (1, 2, 3, 4, 6, 8):
    k.set_param("oneshot_split", pieces)
    for _ in range(3):
        out = k.msm_oneshot_ptr(curve, pts.data_ptr(), n, scp.data_ptr(), n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = k.msm_oneshot_ptr(curve, pts.data_ptr(), n, scp.data_ptr(), n)
    dt = (time.perf_counter() - t0) / 10
    ok = bool((k.to_affine(curve, out) == ref).all())
    for _ in range(3):
        out = k.msm_host_ptr(bases, scp.data_ptr(), n)
    t0 = time.perf_counter()
    for _ in range(10):
        out = k.msm_host_ptr(bases, scp.data_ptr(), n)
    dr = (time.perf_counter() - t0) / 10
    ok = ok and bool((k.to_affine(curve, out) == ref).all())
    print(f"2^{logn} curve {curve} pieces {pieces}: oneshot {dt*1e3:.3f} ms = {n/dt/1e6:.1f} Mpoints/s | registered bases {dr*1e3:.3f} ms = {n/dr/1e6:.1f} Mpoints/s ok={ok}", flush=True)
