"""Development probe: window-collapsed (precomputed) mode vs normal mode."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
CURVE = int(sys.argv[2]) if len(sys.argv) > 2 else k.BN254_G1
CS = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
for logn in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20").split(",")]:
    n = 1 << logn
    bases = k.Bases.generate(CURVE, n, seed=3)
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    torch.cuda.synchronize()
    ref = k.to_affine(CURVE, k.msm_device(bases, d_sc.data_ptr(), n))
    ms, sh = k.last_timing(0)
    print(f"2^{logn} normal   : {ms['total']:.3f} ms", sh, flush=True)
    for c in ([0] + (CS if CS else [17, 18, 19, 20, 21] if logn <= 20 else [20, 21, 22, 23])):
        t0 = time.time()
        bases.precompute(c)
        pre_s = time.time() - t0
        best = None
        for _ in range(3):
            out = k.msm_device(bases, d_sc.data_ptr(), n)
            ms, sh = k.last_timing(0)
            if best is None or ms["total"] < best["total"]:
                best = ms
        ok = bool((k.to_affine(CURVE, out) == ref).all())
        print(f"2^{logn} collapsed c={sh['c']} W={sh['W']} L={sh['L']}: total {best['total']:.3f} acc {best['accumulate']:.3f} reduce {best['reduce']:.3f} fill {best['fill']:.3f} count {best['count']:.3f} fixup {best['fixup']:.3f} host {best['host_finish']:.3f} | precompute {pre_s:.2f} s ok={ok}", flush=True)
    bases.free()
