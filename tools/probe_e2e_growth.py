"""e2e (kgr_msm_oneshot / kgr_msm from pinned host memory) against the number of streamed pieces and the growth factor of their sizes:
python tools/probe_e2e_growth.py [logn] [curve]   (profiles/r02_e2e.md)"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << logn
reps = 10 if logn <= 22 else 4
bases = k.Bases.generate(curve, n, seed=3)
pts = torch.from_numpy(bases.download().view(np.int64)).pin_memory()
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
scp = torch.from_numpy(sc.view(np.int64)).pin_memory()
ref = k.to_affine(curve, k.msm_curve_addition(bases, sc))


def best(fn):
    for _ in range(2):
        out = fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts), out


for pieces in (1, 2, 3, 4, 5, 6):
    for growth in ((100,) if pieces == 1 else (100, 130, 170, 220, 300)):
        k.set_param("oneshot_split", pieces)
        k.set_param("oneshot_growth", growth)
        b1, m1, out = best(lambda: k.msm_oneshot_ptr(curve, pts.data_ptr(), n, scp.data_ptr(), n))
        ok = bool((k.to_affine(curve, out) == ref).all())
        b2, m2, out = best(lambda: k.msm_host_ptr(bases, scp.data_ptr(), n))
        ok = ok and bool((k.to_affine(curve, out) == ref).all())
        print(f"2^{logn} curve {curve} pieces {pieces} growth {growth}: oneshot best {b1*1e3:.3f} mean {m1*1e3:.3f} ms | registered bases best {b2*1e3:.3f} mean {m2*1e3:.3f} ms ok={ok}", flush=True)
k.set_param("oneshot_split", 0)
k.set_param("oneshot_growth", 0)
b1, m1, out = best(lambda: k.msm_oneshot_ptr(curve, pts.data_ptr(), n, scp.data_ptr(), n))
b2, m2, out = best(lambda: k.msm_host_ptr(bases, scp.data_ptr(), n))
print(f"2^{logn} curve {curve} automatic: oneshot best {b1*1e3:.3f} mean {m1*1e3:.3f} ms | registered bases best {b2*1e3:.3f} mean {m2*1e3:.3f} ms", flush=True)
