"""One process, N GPUs (kgr_init with several devices): a single MSM call sharded over the GPUs of the box, the way a prover process would use
the library.  python tools/probe_inprocess_multi.py <log2 n total> [n_gpus]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ng = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
n = 1 << logn
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
scp = torch.from_numpy(sc.view(np.int64)).pin_memory()
res = {}
for g in sorted({1, ng}):
    k.init(list(range(g)))
    bases = k.Bases.generate(k.BN254_G1, n, seed=3)
    for _ in range(2):
        out = k.msm_host_ptr(bases, scp.data_ptr(), n)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        out = k.msm_host_ptr(bases, scp.data_ptr(), n)
    dt = (time.perf_counter() - t0) / reps
    res[g] = k.to_affine(k.BN254_G1, out)
    print(f"2^{logn} points, {g} GPU(s) in one process, host scalars: {dt*1e3:.2f} ms = {n/dt/1e6:.1f} Mpoints/s", flush=True)
    bases.free()
vals = list(res.values())
print("same group element on every device count:", all((v == vals[0]).all() for v in vals))
