"""Host-buffer calls from ordinary (pageable) memory against pinned memory: python tools/probe_pageable.py [logn]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import kogarashi_b200 as k
k.init([0])
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
bases = k.Bases.generate(0, n, seed=3)
pts_np = bases.download()
rng = np.random.default_rng(1)
sc_np = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
pts_pin = torch.from_numpy(pts_np.view(np.int64)).pin_memory()
sc_pin = torch.from_numpy(sc_np.view(np.int64)).pin_memory()
def T(fn, reps=10):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
print(f"2^{logn} oneshot   pinned   {T(lambda: k.msm_oneshot_ptr(0, pts_pin.data_ptr(), n, sc_pin.data_ptr(), n)):.2f} ms | pageable {T(lambda: k.msm_oneshot_ptr(0, pts_np.ctypes.data, n, sc_np.ctypes.data, n)):.2f} ms")
print(f"2^{logn} registered pinned  {T(lambda: k.msm_host_ptr(bases, sc_pin.data_ptr(), n)):.2f} ms | pageable {T(lambda: k.msm_host_ptr(bases, sc_np.ctypes.data, n)):.2f} ms")
