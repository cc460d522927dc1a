import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = k.BN254_G1
for kv in sys.argv[2:]:
    name, v = kv.split("=")
    if name == "curve":
        curve = int(v)
    else:
        k.set_param(name, int(v))
n = 1 << logn
bases = k.Bases.generate(curve, n, seed=3)
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
torch.cuda.synchronize()
for _ in range(3):
    k.msm_device(bases, d_sc.data_ptr(), n)
print(k.last_timing(0))
