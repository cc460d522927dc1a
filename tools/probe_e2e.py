import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
n = 1 << 20
bases = k.Bases.generate(k.BN254_G1, n, seed=3)
pts = torch.from_numpy(bases.download().view(np.int64)).pin_memory()
rng = np.random.default_rng(1)
sc = torch.from_numpy(rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64).view(np.int64)).pin_memory()
for it in range(6):
    t0 = time.perf_counter()
    k.msm_oneshot_ptr(k.BN254_G1, pts.data_ptr(), n, sc.data_ptr(), n)
    wall = (time.perf_counter() - t0) * 1e3
    ms, sh = k.last_timing(0)
    print("oneshot wall %.3f" % wall, {a: round(b, 3) for a, b in ms.items()}, flush=True)
for it in range(3):
    t0 = time.perf_counter()
    k.msm_host_ptr(bases, sc.data_ptr(), n)
    wall = (time.perf_counter() - t0) * 1e3
    ms, sh = k.last_timing(0)
    print("registered wall %.3f" % wall, {a: round(b, 3) for a, b in ms.items()}, flush=True)
