"""Development probe run on the GPU box: integer-pipe microbenchmarks + MSM phase timings.
Writes gpurun_out/probe_<tag>.json.  Not part of the product or the tests."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k  # noqa: E402
from oracle import oracle as A  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 20, 22]
    os.makedirs("gpurun_out", exist_ok=True)
    k.init([0])
    out = {"microbench": k.microbench(), "msm": []}
    print(json.dumps(out["microbench"]), flush=True)
    import torch
    for curve in (k.BN254_G1,):
        for logn in sizes:
            n = 1 << logn
            t0 = time.time()
            bases = k.Bases.generate(curve, n, seed=3)
            gen_s = time.time() - t0
            sc = A.random_field(A.SCALAR_FIELD[curve], min(n, 1 << 20), seed=bytes(range(16)))
            sc = np.tile(sc, (n // sc.shape[0], 1))
            d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
            torch.cuda.synchronize()
            for cbits in ([0] if logn < 20 else [0, 14, 15, 16, 17, 18, 19, 20]):
                if cbits and cbits > logn:
                    continue
                k.set_param("window_bits", cbits)
                best = None
                for it in range(4):
                    k.msm_device(bases, d_sc.data_ptr(), n)
                    ms, shape = k.last_timing(0)
                    if best is None or ms["total"] < best[0]["total"]:
                        best = (ms, shape)
                rec = {"curve": curve, "logn": logn, "gen_s": gen_s, "ms": best[0], "shape": best[1],
                       "mpoints_s": n / best[0]["total"] / 1e3}
                out["msm"].append(rec)
                print(json.dumps(rec), flush=True)
            k.set_param("window_bits", 0)
            # host-scalar path (H2D of scalars inside)
            t0 = time.time()
            k.msm_curve_addition(bases, sc)
            ms, shape = k.last_timing(0)
            rec = {"curve": curve, "logn": logn, "host_scalars_wall_ms": (time.time() - t0) * 1e3, "ms": ms}
            out["msm"].append(rec)
            print(json.dumps(rec), flush=True)
            bases.free()
            del d_sc
    json.dump(out, open(f"gpurun_out/probe_{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
