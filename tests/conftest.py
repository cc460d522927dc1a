import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "msm_vectors.npz"))


@pytest.fixture(scope="session")
def golden_g2():
    return np.load(os.path.join(ROOT, "tests", "golden", "g2_vectors.npz"))


def golden_g2_case_names():
    z = np.load(os.path.join(ROOT, "tests", "golden", "g2_vectors.npz"))
    return sorted({k[: -len("_aff")] for k in z.files if k.endswith("_aff")})


def golden_case_names(prefix=None):
    z = np.load(os.path.join(ROOT, "tests", "golden", "msm_vectors.npz"))
    names = sorted({k[: -len("_aff")] for k in z.files if k.endswith("_aff")})
    return [n for n in names if prefix is None or n.startswith(prefix)]


def same_affine(a, b):
    """zkstd affine equality (macros/curve/weierstrass/group.rs:5-13): identities compare equal regardless of coordinates."""
    a, b = np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64)
    assert a.shape == b.shape  # x, y, then the is_infinity word: (9,), or (17,) for G2
    if int(a[-1]) or int(b[-1]):
        return bool(int(a[-1]) and int(b[-1]))
    return bool((a[:-1] == b[:-1]).all())


def alt_bn128_kat():
    """tests/golden/alt_bn128_kat.json as Montgomery limb arrays: external known-answer vectors (EIP-196 / go-ethereum precompile tests)."""
    import json

    from oracle import pyref as B
    raw = json.load(open(os.path.join(ROOT, "tests", "golden", "alt_bn128_kat.json")))

    def pt(xy):  # canonical hex -> (8,) uint64 x || y Montgomery limbs over Fq
        return np.array([w for v in xy for w in B.int_to_limbs(B.to_mont(int(v, 16), B.FQ))], dtype=np.uint64)

    adds = [(c["name"], pt(c["a"]), pt(c["b"]), pt(c["sum"])) for c in raw["add"]]
    muls = [(c["name"], pt(c["p"]), int(c["k"], 16), pt(c["product"])) for c in raw["mul"]]
    return raw, adds, muls
