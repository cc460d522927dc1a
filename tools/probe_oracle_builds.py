"""Which oracle build is faster on this host: portable (-O3) or -march=native?  MSM of 2^16 / 2^18 points on all threads, best of 3."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as A
u64p = ctypes.POINTER(ctypes.c_uint64)
native = A._native_path()
print("cpu count", os.cpu_count(), "native build:", native)
for path in (A._LIB_PATH, native):
    if not path:
        continue
    L = ctypes.CDLL(path)
    for logn in (16, 18):
        n = 1 << logn
        threads = os.cpu_count()
        k = np.zeros((n, 4), dtype=np.uint64); xy = np.zeros((n, 8), dtype=np.uint64); out = np.zeros(12, dtype=np.uint64)
        L.zko_bench_scalars(0, ctypes.c_uint64(1), ctypes.c_uint64(0), ctypes.c_size_t(n), k.ctypes.data_as(u64p))
        L.zko_fixed_base(0, k.ctypes.data_as(u64p), ctypes.c_size_t(n), threads, xy.ctypes.data_as(u64p))
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            L.zko_msm(0, xy.ctypes.data_as(u64p), None, ctypes.c_size_t(n), k.ctypes.data_as(u64p), ctypes.c_size_t(n), threads, out.ctypes.data_as(u64p))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(os.path.basename(path), f"2^{logn}: {best:.3f} s = {n / best / 1e6:.3f} Mpoints/s", flush=True)
