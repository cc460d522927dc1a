"""Window size x batched-affine levels sweep (round 2): python tools/sweep_c_levels.py <logn> <c_lo> <c_hi> <levels,...>"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
logn, lo, hi = (int(a) for a in sys.argv[1:4])
levels = [int(x) for x in sys.argv[4].split(",")]
n = 1 << logn
bases = k.Bases.generate(0, n, seed=3)
rng = np.random.default_rng(3)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
for lv in levels:
    k.set_param("affine_levels", lv)
    row = []
    for c in range(lo, hi + 1):
        k.set_param("window_bits", c)
        best = None
        for _ in range(3):
            k.msm_device(bases, d_sc.data_ptr(), n)
            t, sh = k.last_timing()
            if best is None or t["total"] < best[0]:
                best = (t["total"], t, sh)
        row.append(f"c={c}:{best[0]:.3f}[sort {best[1]['count'] + best[1]['scan'] + best[1]['fill']:.2f} acc {best[1]['accumulate']:.2f} red {best[1]['reduce']:.2f}]")
    print(f"2^{logn} levels={lv}:", " ".join(row), flush=True)
