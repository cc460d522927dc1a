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
print("sanitize run results ok:", ok)
