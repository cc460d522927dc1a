import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
for logn in (20, 22):
    n = 1 << logn
    bases = k.Bases.generate(k.BN254_G1, n, seed=3)
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    torch.cuda.synchronize()
    ref = None
    for split in (0, 1):
        for rounds in (0, 1, 2, 3):
            for chunk in (0, 96, 128, 160, 256):
                if rounds == 0 and (split or chunk):
                    continue
                k.set_param("affine_split", split); k.set_param("affine_rounds", rounds); k.set_param("chunk", chunk)
                best = None
                for _ in range(3):
                    out = k.msm_device(bases, d_sc.data_ptr(), n)
                    ms, sh = k.last_timing(0)
                    if best is None or ms["total"] < best["total"]:
                        best = ms
                aff = k.to_affine(k.BN254_G1, out)
                if ref is None:
                    ref = aff
                print(f"2^{logn} split={split} rounds={rounds} L={sh['L']}: total {best['total']:.3f} acc {best['accumulate']:.3f} ok={bool((aff == ref).all())}", flush=True)
    k.set_param("affine_rounds", 0); k.set_param("chunk", 0)
    bases.free()
