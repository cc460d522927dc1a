"""Where the Groth16 prover's wall time goes (development probe): python tools/probe_prover.py [log2 constraints]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
from kogarashi_b200 import msm as M
from kogarashi_b200.fft import Fft
from kogarashi_b200.groth16 import Groth16Prover, _canonical
from oracle import groth16_ref as G, pyref as B, oracle as A

k.init([0])
logm = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cs, _ = G.chain_circuit(((1 << logm) - 1) // 3, 3)
E, trap, uvw = G.crs_exponents(cs, B.XorShift128(A.DEFAULT_SEED))
mont = lambda vals: np.array([B.int_to_limbs(B.to_mont(v, B.FR)) for v in vals], dtype=np.uint64).reshape(-1, 4)
pts = lambda c, exps: (M.fixed_base_mul(c, mont(exps)), np.array([1 if e == 0 else 0 for e in exps], dtype=np.uint8))
vk = pts(0, [trap["delta"], trap["alpha"], trap["beta"]])[0]
vk2 = pts(2, [trap["delta"], trap["beta"]])[0]
crs = [pts(0, E[n]) for n in ("a", "b_g1", "h", "l")]
prover = Groth16Prover(vk[0], vk[1], vk[2], *crs[0], *crs[1], *crs[2], *crs[3], vk2[0], vk2[1], *pts(2, E["b_g1"]))
a_ev, b_ev, c_ev = (mont(v) for v in cs.evaluate())
xs, ws = mont(cs.x), mont(cs.w)
r, s = 12345, 67890
f = Fft(E["k"])
def T(fn, n=5):
    fn(); best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, out
t_h, q = T(lambda: f.h_coefficients(a_ev, b_ev, c_ev))
z, jobs = prover._g1_jobs(q, xs, ws, r, 0)
jobs = [(prover.b_g2, z, 0, 0)] + jobs + [(prover.vk_g2_bases, _canonical([s, 1]), 0, 1)]
t_batch, res = T(lambda: k.msm_batch(jobs))
t_seq, _ = T(lambda: [k.msm_batch([j]) for j in jobs])
t_asm, _ = T(lambda: prover._assemble_g1(res[1:6], r, s))
t_all, _ = T(lambda: prover.prove_from_evaluations(E["k"], a_ev, b_ev, c_ev, xs, ws, r, s))
each = [T(lambda j=j: k.msm_batch([j]))[0] for j in jobs]
t_old, _ = T(lambda: (lambda q2: prover.prove(q2, xs, ws, r, s))(f.h_coefficients(a_ev, b_ev, c_ev)))
print(f"2^{logm}: separate H + batch {t_old:.2f} ms | fused call (prove_from_evaluations) {t_all:.2f} ms")
print(f"2^{logm}: h_coefficients {t_h:.2f} ms | batch of {len(jobs)} {t_batch:.2f} ms (one by one {t_seq:.2f}: {[round(x, 2) for x in each]}) | assemble (1 small MSM + host) {t_asm:.2f} | whole {t_all:.2f}")
