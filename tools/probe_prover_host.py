"""Prove latency against the state of the host cores and the way the lanes are driven: python tools/probe_prover_host.py [log2 constraints]"""
import sys, os, time
import numpy as np
sys.path.insert(0, ".")
import kogarashi_b200 as k
from kogarashi_b200 import msm as M
from kogarashi_b200.groth16 import Groth16Prover
from oracle import groth16_ref as G, pyref as B, oracle as A
k.init([0])
logm = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cs, _ = G.chain_circuit(((1 << logm) - 1) // 3, 3)
E, trap, uvw = G.crs_exponents(cs, B.XorShift128(A.DEFAULT_SEED))
mont = lambda vals: np.array([B.int_to_limbs(B.to_mont(v, B.FR)) for v in vals], dtype=np.uint64).reshape(-1, 4)
pts = lambda c, exps: (M.fixed_base_mul(c, mont(exps)), np.array([1 if e == 0 else 0 for e in exps], dtype=np.uint8))
vk = pts(0, [trap["delta"], trap["alpha"], trap["beta"]])[0]; vk2 = pts(2, [trap["delta"], trap["beta"]])[0]
crs = [pts(0, E[n]) for n in ("a", "b_g1", "h", "l")]; crs2 = pts(2, E["b_g1"])
a_ev, b_ev, c_ev = (mont(v) for v in cs.evaluate()); xs, ws = mont(cs.x), mont(cs.w)
p = Groth16Prover(vk[0], vk[1], vk[2], *crs[0], *crs[1], *crs[2], *crs[3], vk2[0], vk2[1], *crs2)
burn_pts = A.random_points(0, 1 << 12); burn_sc = A.random_field(A.FIELD_FR, 1 << 16)
def measure(n=20):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); p.prove_from_evaluations(E["k"], a_ev, b_ev, c_ev, xs, ws, 5, 7); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts[-10:])
for threads in (1, 0, 1, 0):
    k.set_param("lane_threads", threads)
    measure(5)
    time.sleep(1.0)
    idle = measure()
    for _ in range(3):
        A.msm(0, np.tile(burn_pts, (16, 1)), burn_sc, threads=os.cpu_count())
    busy = measure(10)
    print(f"2^{logm} lane_threads={threads}: host idle before the calls {idle:.2f} ms | right after multi-threaded CPU work {busy:.2f} ms", flush=True)
