"""Nova-shaped scalars (half zero, a quarter equal, a quarter uniform) against uniform ones: python tools/probe_skew.py [logn ...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
for logn in [int(a) for a in sys.argv[1:]] or [15, 20]:
    n = 1 << logn
    bases = k.Bases.generate(0, n, seed=3)
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    skew = sc.copy()
    skew[: n // 2] = 0
    skew[n // 2: 3 * n // 4] = skew[n // 2]
    small = np.zeros_like(sc)
    small[:, 0] = rng.integers(0, 2, size=n, dtype=np.uint64)          # 0 / 1 witnesses (canonical format)
    for name, s, fmt in (("uniform", sc, k.SCALARS_MONTGOMERY), ("half zero, quarter equal", skew, k.SCALARS_MONTGOMERY), ("0/1 canonical", small, k.SCALARS_CANONICAL)):
        d = torch.from_numpy(s.view(np.int64)).cuda()
        best = None
        for _ in range(4):
            k.msm_device(bases, d.data_ptr(), n, scalar_fmt=fmt)
            t, sh = k.last_timing(0)
            if best is None or t["total"] < best[0]["total"]:
                best = (t, sh)
        t, sh = best
        print(f"2^{logn} {name}: total {t['total']:.3f} sort {t['count'] + t['scan'] + t['fill']:.3f} acc {t['accumulate']:.3f} fixup {t['fixup']:.3f} reduce {t['reduce']:.3f} host {t['host_finish']:.3f} c={sh['c']} L={sh['L']}", flush=True)
    bases.free()
