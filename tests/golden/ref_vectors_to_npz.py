"""Turn the JSON lines printed by tests/golden/gen_from_reference.rs (run inside the reference, see oracle/_ref/README.md) into
tests/golden/msm_vectors_ref.npz: for every case `<name>_pts` (n, 8) uint64, `<name>_inf` (n,) uint8, `<name>_sc` (m, 4) uint64 and
`<name>_aff` (9,) uint64 — the layout of tests/golden/msm_vectors.npz."""
import json
import sys

import numpy as np


def limbs(hex4):
    return [int(h, 16) for h in hex4]


def main(src, dst):
    out = {}
    for line in open(src):
        line = line.strip()
        if not line.startswith("{"):
            continue
        c = json.loads(line)
        name = c["name"]
        pts = np.array([limbs(p[0]) + limbs(p[1]) for p in c["points"]], dtype=np.uint64).reshape(-1, 8)
        out[name + "_pts"] = pts
        out[name + "_inf"] = np.array([p[2] for p in c["points"]], dtype=np.uint8)
        out[name + "_sc"] = np.array([limbs(s) for s in c["scalars"]], dtype=np.uint64).reshape(-1, 4)
        out[name + "_aff"] = np.array(limbs(c["affine"][0]) + limbs(c["affine"][1]) + [c["affine"][2]], dtype=np.uint64)
    np.savez_compressed(dst, **out)
    print(f"{len(out) // 4} cases -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
