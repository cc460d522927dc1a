import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
for logn in (18, 20, 22):
    n = 1 << logn
    bases = k.Bases.generate(k.BN254_G1, n, seed=3)
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    torch.cuda.synchronize()
    for mode in (0, 1, 2, 3, 4):
        k.set_param("affine_rounds", mode)
        best = None
        for _ in range(3):
            k.msm_device(bases, d_sc.data_ptr(), n)
            ms, sh = k.last_timing(0)
            if best is None or ms["total"] < best["total"]:
                best = ms
        print(f"2^{logn} affine_rounds={mode}: total {best['total']:.3f} L={sh['L']} acc {best['accumulate']:.3f} c={sh['c']}", flush=True)
    k.set_param('affine_rounds', 0)
    bases.free()
