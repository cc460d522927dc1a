"""Development sweep of the reduce parameters at 2^20 (not part of the product)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
bases = k.Bases.generate(k.BN254_G1, n, seed=3)
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
torch.cuda.synchronize()
for c in (15, 16, 17):
    for K in (8, 16, 32, 64):
        for stop in (1024, 4096, 16384):
            k.set_param("window_bits", c); k.set_param("reduce_fanin", K); k.set_param("running_sum_stop", stop)
            best = None
            for _ in range(3):
                k.msm_device(bases, d_sc.data_ptr(), n)
                ms, sh = k.last_timing(0)
                if best is None or ms["total"] < best["total"]:
                    best = ms
            print(f"c={c} K={K} stop={stop}: total {best['total']:.3f} acc {best['accumulate']:.3f} reduce {best['reduce']:.3f} fixup {best['fixup']:.3f}", flush=True)
