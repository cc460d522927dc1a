"""Window-size sweep at small n (development probe): python tools/sweep_small.py [curve]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
curve = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rng = np.random.default_rng(3)
for logn in (8, 10, 12, 13, 14, 15, 16, 17, 18):
    n = 1 << logn
    bases = k.Bases.generate(curve, n, seed=3)
    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    row = []
    for c in [0] + list(range(6, 18)):
        k.set_param("window_bits", c)
        best = None
        for _ in range(4):
            k.msm_curve_addition(bases, sc, scalar_fmt=k.SCALARS_CANONICAL)
            t, sh = k.last_timing()
            v = t["total"] - t["h2d"]
            if best is None or v < best[0]:
                best = (v, t, sh)
        row.append(f"c={best[2]['c']}:{best[0]:.3f}[{best[1]['accumulate']:.2f}/{best[1]['fixup']:.2f}/{best[1]['reduce']:.2f}]" + (f"(auto; acc {best[1]['accumulate']:.2f} fix {best[1]['fixup']:.2f} red {best[1]['reduce']:.2f} host {best[1]['host_finish']:.2f})" if c == 0 else ""))
    k.set_param("window_bits", 0)
    print(f"2^{logn}:", " ".join(row), flush=True)
    bases.free()
