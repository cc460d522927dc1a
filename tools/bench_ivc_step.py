"""Vector work of one Nova IVC step on one curve with everything resident on the GPU:
nova/src/ivc.rs:160-205 -> r1cs_instance_and_witness (commit of the fresh witness, relaxed_r1cs.rs:36), Prover::prove (prover.rs:24-50:
cross term T, commit(T)), RelaxedR1csWitness::fold (witness.rs:56-71: W <- W + r W2, E <- E + r T).  An IVC step does this once per curve
of the cycle: 2 cross terms + 4 commitments + 4 folds.

  python tools/bench_ivc_step.py <constraints_log2 or chain steps> [curve: 0 BN254 G1 (Fr) | 1 Grumpkin (Fq)]

Paths measured on the chained x^3 + x + 5 circuit:
  resident     z1, z2, E, T are kgr_vec_t; only the two 72-byte commitments come back            (round 2)
  resident+w   the same plus the upload of the fresh witness z2 (it is produced on the host)
  host         kgr_nova_cross_term / kgr_pedersen_commit / kgr_vec_fold with host buffers          (round 1)
  cpu          the restated reference loops (oracle) with msm_curve_addition as the commitment; the reference's own commit is a naive
               per-element scalar multiplication (pedersen.rs:15-20), timed on a 2^10 sample and extrapolated
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kogarashi_b200 as k  # noqa: E402
from kogarashi_b200 import nova  # noqa: E402
from nova_util import chain_r1cs, mont  # noqa: E402
from oracle import oracle as A  # noqa: E402
from oracle import pyref as B  # noqa: E402


def best(f, reps=5):
    f()
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    arg = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    steps = ((1 << arg) - 1) // 3 if arg <= 24 else arg
    curve = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    fid = A.SCALAR_FIELD[curve]
    p = B.FQ if fid == A.FIELD_FQ else B.FR
    k.init([0])
    t0 = time.perf_counter()
    m, n_z, mats, z1_int = chain_r1cs(steps, 3, p)
    z2_int = chain_r1cs(steps, 5, p)[3]
    z1, z2 = mont(z1_int, p), mont(z2_int, p)
    setup_s = time.perf_counter() - t0
    l = 3
    shape = nova.R1csShape(fid, m, n_z, *mats)
    ck = k.PedersenCommitment.new_benchmark(curve, max(m, n_z).bit_length(), seed=9)
    r = mont([0x1F2E3D4C5B6A79880123456789ABCDEF % p], p)[0]
    d_z1, d_z2 = nova.DeviceVec(fid, z1), nova.DeviceVec(fid, z2)
    d_e, d_t = nova.DeviceVec(fid, n=m), nova.DeviceVec(fid, n=m)
    res = {}

    def resident(upload):
        if upload:
            d_z2.write(0, z2)
        cw = d_z2.commit(ck, off=l)
        ct = nova.cross_term_device(shape, d_z1, d_z2, t=d_t, ck=ck)
        d_z1.fold(d_z2, r)
        d_e.fold(d_t, r)
        return cw, ct

    res["resident_ms"] = best(lambda: resident(False))
    res["resident_with_witness_upload_ms"] = best(lambda: resident(True))
    # correctness of one step from a fresh state against the oracle
    d_z1.write(0, z1)
    d_e2 = nova.DeviceVec(fid, n=m)
    cw = d_z2.commit(ck, off=l)
    ct = nova.cross_term_device(shape, d_z1, d_z2, t=d_t, ck=ck)
    d_z1.fold(d_z2, r)
    d_e2.fold(d_t, r)
    t_ref = A.cross_term(fid, m, *mats, z1, z2)
    g = ck.g.download()
    ok = bool((d_t.download() == t_ref).all() and (d_z1.download() == A.vec_fold(fid, z1, z2, r)).all()
              and (d_e2.download() == A.vec_fold(fid, np.zeros_like(t_ref), t_ref, r)).all()
              and (ct[:8] == A.to_affine(curve, A.msm(curve, g[:m], t_ref))[:8]).all() and (cw[:8] == A.to_affine(curve, A.msm(curve, g[:n_z - l], z2[l:]))[:8]).all())
    res["bit_exact_with_oracle"] = ok

    e_host = np.zeros((m, 4), dtype=np.uint64)

    def host():
        cw = ck.commit(z2[l:])
        t, ct = shape.cross_term(z1, z2, ck=ck)
        zf = nova.vec_fold(fid, z1, z2, r)
        ef = nova.vec_fold(fid, e_host, t, r)
        return cw, ct, zf, ef

    res["host_buffers_ms"] = best(host, reps=3)

    cores = os.cpu_count() or 1

    def cpu():
        A.msm(curve, g[:n_z - l], z2[l:], threads=cores)
        t = A.cross_term(fid, m, *mats, z1, z2)
        A.msm(curve, g[:m], t, threads=cores)
        A.vec_fold(fid, z1, z2, r)
        A.vec_fold(fid, e_host, t, r)

    t0 = time.perf_counter()
    cpu()
    res["cpu_reference_loops_with_msm_commit_ms"] = (time.perf_counter() - t0) * 1e3
    ns = min(1 << 10, m)
    t0 = time.perf_counter()
    A.pedersen_commit(curve, g[:ns], z2[l:l + ns])
    naive_per = (time.perf_counter() - t0) / ns
    res["cpu_reference_naive_commit_extrapolated_ms"] = naive_per * (m + n_z - l) * 1e3
    rec = {"row": "N4 Nova IVC-step vector work on one curve: commit(W), cross term, commit(T), fold z, fold E", "curve": curve, "constraints": m, "n_z": n_z,
           "cores": cores, "setup_python_s": setup_s, **res,
           "speedup_resident_vs_host_buffers": res["host_buffers_ms"] / res["resident_ms"],
           "speedup_resident_vs_cpu_msm_commit": res["cpu_reference_loops_with_msm_commit_ms"] / res["resident_ms"]}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
