"""G2 MSM timing probe: python tools/probe_g2.py [logn ...]  (device-generated bases, canonical random scalars < 2^252)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import kogarashi_b200 as k  # noqa: E402

k.init()
rng = np.random.default_rng(7)
out = []
for logn in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
    n = 1 << logn
    for curve, name in ((k.BN254_G1, "g1"), (k.BN254_G2, "g2")):
        bases = k.Bases.generate(curve, n, seed=3)
        sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
        sc[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
        best = None
        for _ in range(4):
            k.msm_curve_addition(bases, sc, scalar_fmt=k.SCALARS_CANONICAL)
            t, shape = k.last_timing()
            dev = t["total"] - t["h2d"]
            if best is None or dev < best[0]:
                best = (dev, t, shape)
        rec = dict(curve=name, logn=logn, ms=round(best[0], 3), mpts=round(n / best[0] / 1e3, 1), shape=best[2],
                   phases={a: round(b, 3) for a, b in best[1].items()})
        out.append(rec)
        print(json.dumps(rec), flush=True)
        if curve == k.BN254_G2 and logn <= 20:
            bases.precompute()
            k.msm_curve_addition(bases, sc, scalar_fmt=k.SCALARS_CANONICAL)
            k.msm_curve_addition(bases, sc, scalar_fmt=k.SCALARS_CANONICAL)
            t, shape = k.last_timing()
            print(json.dumps(dict(curve="g2_precomputed", logn=logn, ms=round(t["total"] - t["h2d"], 3), shape=shape)), flush=True)
        bases.free()
json.dump(out, open("gpurun_out/probe_g2.json", "w"))
