"""Window-size sweep at one size (development probe): python tools/sweep_c.py <curve> <logn> <c_lo> <c_hi>"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
curve, logn, lo, hi = (int(a) for a in sys.argv[1:5])
n = 1 << logn
bases = k.Bases.generate(curve, n, seed=3)
rng = np.random.default_rng(3)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
row = []
for c in [0] + list(range(lo, hi + 1)):
    k.set_param("window_bits", c)
    best = None
    for _ in range(4):
        k.msm_device(bases, d_sc.data_ptr(), n)
        t, sh = k.last_timing()
        if best is None or t["total"] < best[0]:
            best = (t["total"], t, sh)
    row.append(f"c={best[2]['c']}:{best[0]:.3f}" + ("(auto)" if c == 0 else "") + f"[acc {best[1]['accumulate']:.2f} red {best[1]['reduce']:.2f} fix {best[1]['fixup']:.2f}]")
print(f"curve {curve} 2^{logn}:", " ".join(row), flush=True)
