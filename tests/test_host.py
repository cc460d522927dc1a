"""CPU tests of the product's host-side logic: the pipeline bodies executed thread by thread on the
CPU (tests/host_emu.cpp), the C-ABI library's symbol table, and its host-only helpers."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
from conftest import ROOT, same_affine

from oracle import oracle as A

EMU = os.path.join(ROOT, "tests", "_build", "host_emu")


@pytest.fixture(scope="module")
def host_emu():
    os.makedirs(os.path.dirname(EMU), exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "kogarashi_b200", "csrc", f) for f in ("field.cuh", "modinv.cuh", "curve.cuh", "msm_kernels.cuh")]
    if not os.path.exists(EMU) or any(os.path.getmtime(d) > os.path.getmtime(EMU) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", EMU, src])
    return EMU


# curve n c L K mode seed   (mode: 0 uniform, 1 skewed zeros/ones/r-1, 2 duplicates + opposites + identity bases, 3 canonical scalars)
# seed bits select variants: bit 0 window-major fill, bits 1-2 batched-affine tree levels in front of the XYZZ accumulation, bit 4 fold reduce, bit 5 two streamed pieces,
# bit 6 chunk length from the entries actually present
EMU_CASES = [
    (0, 1, 4, 16, 16, 0, 1), (0, 2, 1, 16, 2, 0, 2), (0, 3, 3, 16, 4, 0, 3), (0, 37, 5, 8, 4, 0, 4), (0, 300, 8, 16, 16, 0, 5),
    (0, 300, 8, 16, 16, 1, 6), (0, 300, 7, 5, 8, 2, 7), (1, 257, 9, 32, 16, 0, 8), (1, 200, 6, 16, 2, 2, 9), (0, 129, 1, 16, 16, 0, 10),
    (0, 500, 11, 64, 16, 3, 11), (1, 64, 13, 16, 16, 1, 12), (0, 700, 10, 256, 4, 1, 13), (1, 333, 12, 1, 16, 2, 14),
    # running_sum_stop > 1: the weighting pass + tree sum close the reduction; small chunks + skew: hot-bucket worklist
    (0, 300, 8, 16, 4, 0, 15, 16), (1, 300, 9, 2, 2, 1, 16, 64), (0, 400, 6, 3, 4, 2, 17, 4096), (0, 600, 10, 2, 16, 1, 18, 8),
    # mode + 10: window-collapsed mode on a precomputed table (kgr_bases_precompute)
    (0, 200, 7, 16, 16, 10, 19, 4096), (1, 150, 5, 4, 4, 11, 20, 8), (0, 120, 11, 32, 16, 12, 21, 1), (1, 90, 3, 8, 2, 13, 22, 4096),
    # curve 2: BN254 G2 (Fq2 coordinates) through the same bodies
    (2, 1, 4, 16, 16, 0, 32), (2, 150, 7, 16, 4, 0, 33), (2, 200, 6, 5, 8, 2, 34), (2, 180, 8, 8, 16, 1, 35, 16), (2, 100, 5, 4, 4, 12, 36, 8),
    (2, 120, 9, 32, 2, 3, 48),
    # seed bit 4: the fold reduce (default GPU reduction) through the bodies shared with k_fold / k_fold_tail / k_vsum1 / k_fold_combine
    (0, 300, 8, 16, 16, 0, 16), (0, 300, 7, 5, 8, 2, 17), (1, 257, 9, 32, 16, 1, 48), (2, 150, 7, 16, 4, 0, 49), (0, 200, 7, 16, 16, 10, 51, 4096),
    (1, 64, 3, 16, 16, 1, 16), (0, 500, 11, 64, 16, 3, 18), (1, 90, 10, 8, 2, 13, 52, 4096),
    # seed bit 5: the call streamed in two pieces accumulated INTO one bucket set (MsmShape::into), with / without affine levels, fold reduce,
    # window tables, hot buckets (several of the curve-2 cases above have the bit set as well)
    (0, 300, 8, 16, 16, 0, 37), (1, 257, 9, 32, 16, 1, 54), (0, 3, 3, 16, 4, 0, 35), (0, 600, 10, 2, 16, 1, 50, 8), (0, 2, 5, 16, 16, 2, 32),
    (1, 333, 12, 1, 16, 2, 46), (0, 120, 11, 32, 16, 12, 53, 1), (0, 700, 6, 3, 4, 1, 39, 4096),
    # seed bit 6: chunk length derived from the entries actually present (eff_chunk_len): skewed scalars, affine levels (far fewer nodes than the
    # bound), streamed pieces, window tables, G2
    (0, 300, 8, 64, 16, 1, 64), (0, 700, 10, 256, 4, 1, 70 + 16), (1, 257, 9, 32, 16, 1, 64 + 54), (0, 500, 11, 64, 16, 3, 64 + 18), (2, 150, 7, 16, 4, 1, 64 + 49),
    (0, 200, 7, 16, 16, 11, 64 + 51, 4096), (1, 333, 6, 24, 16, 2, 64 + 6), (0, 600, 5, 100, 16, 1, 64 + 4),
]


@pytest.mark.parametrize("case", EMU_CASES)
def test_pipeline_bodies_on_cpu(host_emu, case):
    out = subprocess.run([host_emu] + [str(x) for x in case], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


@pytest.fixture(scope="module")
def host_emu_r1cs():
    exe = os.path.join(ROOT, "tests", "_build", "host_emu_r1cs")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_emu_r1cs.cpp")
    deps = [src] + [os.path.join(ROOT, "kogarashi_b200", "csrc", f) for f in ("field.cuh", "r1cs_kernels.cuh")] + [os.path.join(ROOT, "oracle", "zkstd_oracle.hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, src])
    return exe


@pytest.mark.parametrize("case", [(0, 200, 57, 1), (1, 333, 90, 2), (1, 1, 1, 3), (0, 64, 300, 4)])
def test_r1cs_bodies_on_cpu(host_emu_r1cs, case):
    """Nova folding kernels' per-row bodies (sparse product, fused cross term, fold) executed on the host against the oracle's loops."""
    out = subprocess.run([host_emu_r1cs] + [str(x) for x in case], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_library_exports_every_declared_symbol():
    from kogarashi_b200 import _lib
    _lib.build()
    hdr = open(os.path.join(ROOT, "include", "kgr_msm.h")).read()
    declared = set(re.findall(r"\b(kgr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import kogarashi_b200 as k
    with pytest.raises(k.KgrError):
        k.init()
    with pytest.raises(k.KgrError):
        k.msm_curve_addition(np.zeros((1, 8), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64))


@pytest.mark.parametrize("curve", [0, 1])
def test_host_helpers_match_oracle(curve):
    """kgr_to_affine / kgr_proj_add are host code (they combine per-GPU partial sums); check them on the CPU."""
    import kogarashi_b200 as k
    g = A.generator(curve)
    one = A.field_op(A.BASE_FIELD[curve], "to_mont", np.array([1, 0, 0, 0], dtype=np.uint64))
    gp = np.concatenate([g, one])
    ks = A.random_field(A.SCALAR_FIELD[curve], 6, seed=bytes(range(100, 116)))
    pts = [A.scalar_point(curve, gp, s) for s in ks]        # non-trivial Z
    for p in pts:
        assert same_affine(k.to_affine(curve, p), A.to_affine(curve, p))
    for a, b in zip(pts[:-1], pts[1:]):
        assert same_affine(k.to_affine(curve, k.proj_add(curve, a, b)), A.to_affine(curve, A.point_op(curve, 0, a, b)))
    ident = np.concatenate([np.zeros(4, dtype=np.uint64), one, np.zeros(4, dtype=np.uint64)])
    assert int(k.to_affine(curve, ident)[8]) == 1
    assert same_affine(k.to_affine(curve, k.proj_add(curve, pts[0], ident)), A.to_affine(curve, pts[0]))
    assert same_affine(k.to_affine(curve, k.proj_add(curve, pts[0], pts[0])), A.to_affine(curve, A.point_op(curve, 1, pts[0])))
    neg = pts[0].copy()
    neg[4:8] = A.field_op(A.BASE_FIELD[curve], "neg", pts[0][4:8])
    assert int(k.to_affine(curve, k.proj_add(curve, pts[0], neg))[8]) == 1
