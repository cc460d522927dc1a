"""Round 2: batched-affine tree levels (affine_levels = r) in front of the XYZZ accumulation: same point, per-phase device times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
logs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10, 16, 20, 24]
levels = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
for logn in logs:
    n = 1 << logn
    bases = k.Bases.generate(k.BN254_G1, n, seed=3)
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    skew = sc.copy()
    skew[: n // 2] = 0
    skew[n // 2: 3 * n // 4] = skew[n // 2]
    for name, s in (("uniform", sc), ("skewed", skew)):
        d_sc = torch.from_numpy(s.view(np.int64)).cuda()
        torch.cuda.synchronize()
        ref = None
        for lv in levels:
            k.set_param("affine_levels", lv)
            best = None
            for _ in range(3):
                out = k.to_affine(k.BN254_G1, k.msm_device(bases, d_sc.data_ptr(), n))
                ms, sh = k.last_timing(0)
                if best is None or ms["total"] < best["total"]:
                    best = ms
            if ref is None:
                ref = out
            ok = bool((out == ref).all())
            print(f"2^{logn} {name:8s} affine_levels={lv}: total {best['total']:.3f} sort {best['count'] + best['scan'] + best['fill']:.3f} "
                  f"acc {best['accumulate']:.3f} fixup {best['fixup']:.3f} reduce {best['reduce']:.3f} c={sh['c']} L={sh['L']} same_point={ok}", flush=True)
            assert ok
    k.set_param("affine_levels", 0)
    bases.free()
