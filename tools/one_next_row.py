"""One invocation of a next-row kernel for ncu captures: python tools/one_next_row.py ntt [log2 n] | nova [chain steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import kogarashi_b200 as k
k.init([0])
what = sys.argv[1]
rng = np.random.default_rng(1)
if what == "ntt":
    from kogarashi_b200.fft import Fft
    kk = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    v = rng.integers(0, 1 << 62, size=(1 << kk, 4), dtype=np.uint64)
    for _ in range(2):
        Fft(kk).dft(v)
else:
    from nova_util import chain_r1cs
    from oracle import pyref as B
    from kogarashi_b200 import nova
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 349525
    m, n_z, mats, z1 = chain_r1cs(steps, 3, B.FQ)
    R256 = 1 << 256
    to_m = lambda vals: np.frombuffer(b"".join((x * R256 % B.FQ).to_bytes(32, "little") for x in vals), dtype=np.uint64).reshape(-1, 4).copy()
    z = to_m(z1)
    shape = nova.R1csShape(0, m, n_z, *mats)
    for _ in range(2):
        shape.cross_term(z, z, want_t=False)
