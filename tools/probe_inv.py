import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
from kogarashi_b200 import msm as M
k.init([0])
n = 1 << 20
rng = np.random.default_rng(1)
a = rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
for op, name in ((2, "mul"), (7, "fermat inv"), (9, "safegcd inv")):
    M.test_field_op(0, op, a, a)
    t0 = time.perf_counter(); M.test_field_op(0, op, a, a); dt = time.perf_counter() - t0
    print(name, "wall incl copies %.2f ms" % (dt * 1e3), flush=True)
