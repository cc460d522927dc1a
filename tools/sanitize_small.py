"""Small MSMs on all three curves for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kogarashi_b200 as k
from oracle import oracle as A
k.init([0])
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "msm_vectors.npz"))
ok = True
for name in ("g1_uniform_1024", "gr_skewed_128", "g1_dup_neg_96", "gr_identity_bases_40", "g1_uniform_3", "gr_uniform_0"):
    curve = 0 if name.startswith("g1_") else 1
    for chunk in (0, 2):
        k.set_param("chunk", chunk)
        got = k.to_affine(curve, k.msm_curve_addition(z[name + "_pts"], z[name + "_sc"], curve=curve, inf=z[name + "_inf"]))
        ok &= bool((got[:8] == z[name + "_aff"][:8]).all() or (got[8] and z[name + "_aff"][8]))
    b = k.Bases(curve, z[name + "_pts"], z[name + "_inf"]).precompute(5)
    got = k.to_affine(curve, k.msm_curve_addition(b, z[name + "_sc"]))
    ok &= bool((got[:8] == z[name + "_aff"][:8]).all() or (got[8] and z[name + "_aff"][8]))
    b.free()
# round 2 kernels, forced at a small size: radix-partition sort, batched-affine levels (3: odd and even levels, padding-free offsets), a host-buffer
# call streamed in two pieces (per-piece buckets + merge); the quad-cooperative fix-up / reduction tail runs in every case of this script
for name in ("g1_uniform_1024", "gr_skewed_128", "g1_dup_neg_96"):
    curve = 0 if name.startswith("g1_") else 1
    for pname, v in (("sort_mode", 2), ("affine_levels", 3), ("oneshot_split", 2), ("chunk", 0)):
        k.set_param(pname, v)
    got = k.to_affine(curve, k.msm_curve_addition(z[name + "_pts"], z[name + "_sc"], curve=curve, inf=z[name + "_inf"]))
    ok &= bool((got[:8] == z[name + "_aff"][:8]).all() or (got[8] and z[name + "_aff"][8]))
    for pname, v in (("sort_mode", -1), ("affine_levels", -1), ("oneshot_split", 0)):
        k.set_param(pname, v)
# a hot bucket cut into thousands of pieces (all scalars equal, 2-entry chunks): k_fixup_long sums it on several CTAs (slice sums + arrival
# counter), then with the chunk length derived from the entries on the device (chunk = 0)
pts_h = A.random_points(0, 4096, seed=bytes(range(7, 23)))
sc_h = np.repeat(A.random_field(A.FIELD_FR, 1, seed=bytes(range(9, 25))), 4096, axis=0)
exp_h = A.to_affine(0, A.msm(0, pts_h, sc_h))
for chunk in (2, 0):
    k.set_param("chunk", chunk)
    got = k.to_affine(0, k.msm_curve_addition(pts_h, sc_h, curve=0))
    ok &= bool((got == exp_h).all())
k.set_param("chunk", 0)
z2 = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "g2_vectors.npz"))
for name in ("g2_uniform_128", "g2_dup_neg_48", "g2_identity_bases_24", "g2_skewed_64", "g2_uniform_0"):
    for chunk in (0, 2):
        k.set_param("chunk", chunk)
        got = k.to_affine(2, k.msm_curve_addition(z2[name + "_pts"], z2[name + "_sc"], curve=2, inf=z2[name + "_inf"]))
        ok &= bool((got[:16] == z2[name + "_aff"][:16]).all() or (got[16] and z2[name + "_aff"][16]))
    b = k.Bases(2, z2[name + "_pts"], z2[name + "_inf"]).precompute(5)
    got = k.to_affine(2, k.msm_curve_addition(b, z2[name + "_sc"]))
    ok &= bool((got[:16] == z2[name + "_aff"][:16]).all() or (got[16] and z2[name + "_aff"][16]))
    b.free()
# batch lanes (several streams + host threads) and the R1CS kernels
b1 = k.Bases(0, z["g1_uniform_1024_pts"], z["g1_uniform_1024_inf"])
b2 = k.Bases(2, z2["g2_uniform_128_pts"], z2["g2_uniform_128_inf"])
got = k.msm_batch([(b2, z2["g2_uniform_128_sc"]), (b1, z["g1_uniform_1024_sc"]), (b1, z["g1_uniform_1024_sc"][:500], 100)])
ok &= bool((k.to_affine(0, got[1])[:8] == z["g1_uniform_1024_aff"][:8]).all() and (k.to_affine(2, got[0])[:16] == z2["g2_uniform_128_aff"][:16]).all())
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nova_util import chain_r1cs, mont
from oracle import pyref as B
from kogarashi_b200 import nova
m, n_z, mats, z1 = chain_r1cs(100, 3, B.FQ)
zz2 = chain_r1cs(100, 4, B.FQ)[3]
shape = nova.R1csShape(0, m, n_z, *mats)
ck = k.PedersenCommitment(1, A.random_points(1, 512))
t, commit = shape.cross_term(mont(z1, B.FQ), mont(zz2, B.FQ), ck=ck)
ok &= bool((t == A.cross_term(0, m, *mats, mont(z1, B.FQ), mont(zz2, B.FQ))).all())
ok &= bool((nova.vec_fold(0, t, t, t[0]) == A.vec_fold(0, t, t, t[0])).all())
fb = k.Bases.generate(0, 300, seed=5)
print("sanitize run results ok:", ok)
