"""Round 2: radix-partition sort (sort_mode = 2) against the counting sort (0 / 1): same group element, per-phase device times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
import torch
k.init([0])
logs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [12, 16, 18, 20, 22, 24]
for logn in logs:
    n = 1 << logn
    bases = k.Bases.generate(k.BN254_G1, n, seed=3)
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    skew = sc.copy()
    skew[: n // 2] = 0
    skew[n // 2: 3 * n // 4, 1:] = 0
    skew[n // 2: 3 * n // 4, 0] = 1   # canonical 1 after from_mont? (Montgomery input: value 1 * R^-1) -- just a repeated scalar
    for name, s in (("uniform", sc), ("skewed", skew)):
        d_sc = torch.from_numpy(s.view(np.int64)).cuda()
        torch.cuda.synchronize()
        ref = None
        for mode in (0, 1, 2):
            k.set_param("sort_mode", mode)
            best = None
            for _ in range(3):
                out = k.to_affine(k.BN254_G1, k.msm_device(bases, d_sc.data_ptr(), n))
                ms, sh = k.last_timing(0)
                if best is None or ms["total"] < best["total"]:
                    best = ms
            if ref is None:
                ref = out
            ok = bool((out == ref).all())
            print(f"2^{logn} {name:8s} sort_mode={mode}: total {best['total']:.3f} count {best['count']:.3f} scan {best['scan']:.3f} fill {best['fill']:.3f} "
                  f"acc {best['accumulate']:.3f} fixup {best['fixup']:.3f} reduce {best['reduce']:.3f} c={sh['c']} same_point={ok}", flush=True)
            assert ok
    if logn <= 20:
        # window-collapsed table mode shares one bucket set between the windows (nseg = 1)
        k.set_param("sort_mode", 0)
        bases.precompute(0)
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        ref = None
        for mode in (0, 2):
            k.set_param("sort_mode", mode)
            out = k.to_affine(k.BN254_G1, k.msm_device(bases, d_sc.data_ptr(), n))
            out = k.to_affine(k.BN254_G1, k.msm_device(bases, d_sc.data_ptr(), n))
            ms, sh = k.last_timing(0)
            if ref is None:
                ref = out
            print(f"2^{logn} table    sort_mode={mode}: total {ms['total']:.3f} count {ms['count']:.3f} fill {ms['fill']:.3f} acc {ms['accumulate']:.3f} c={sh['c']} same_point={bool((out == ref).all())}", flush=True)
            assert (out == ref).all()
    k.set_param("sort_mode", -1)
    bases.free()
