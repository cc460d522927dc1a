"""Development probe: NTT device time vs size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
for logn in (16, 18, 20, 22, 24):
    n = 1 << logn
    rng = np.random.default_rng(1)
    x = rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
    d = torch.from_numpy(x.view(np.int64)).cuda()
    f = k.Fft(logn)
    torch.cuda.synchronize()
    for op in ("dft", "coset_idft"):
        best = 1e9
        for _ in range(5):
            f.transform_device(op, d.data_ptr())
            best = min(best, k.last_timing(0)[0]["total"])
        muls = n // 2 * logn
        print(f"2^{logn} {op}: {best:.3f} ms  {n / best / 1e3:.1f} Melem/s  butterflies {muls / best / 1e6:.1f} G/s  bytes moved/algorithmic(64B*n) {64 * n / best / 1e6:.1f} GB/s", flush=True)
