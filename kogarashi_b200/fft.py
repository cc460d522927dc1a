"""Mirror of the reference's radix-2 FFT over bn254 Fr on the GPU.

Reference (groth16/src/fft.rs): `Fft::<Fr>::new(k)` then `dft / idft / coset_dft / coset_idft` on
`Coefficients` / `PointsValue` vectors (:92-127); inputs shorter than 2^k are zero padded (`prepare_fft`, :157-162) and
the inverse transforms return `Coefficients::new(..)`, which drops trailing zeros (groth16/src/poly.rs:61-63).
Arrays are (len, 4) uint64 Montgomery limbs (`Fr.0`).
"""
import ctypes

import numpy as np

from . import _lib
from .msm import _c, _u64

OPS = dict(dft=0, idft=1, coset_dft=2, coset_idft=3)


class Fft:
    def __init__(self, k):
        assert 1 <= k <= 28  # fft.rs:28 asserts k >= 1; Fr has 2-adicity S = 28 (bn254/src/fr.rs:53)
        _lib.ensure_init()
        self.k, self.n = k, 1 << k

    def _run(self, op, values):
        v = _c(values).reshape(-1, 4)
        assert v.shape[0] <= self.n
        out = np.zeros((self.n, 4), dtype=np.uint64)
        n_out = ctypes.c_size_t()
        _lib.check(_lib.lib().kgr_ntt(self.k, OPS[op], _u64(v), v.shape[0], _u64(out), ctypes.byref(n_out)))
        return out[: n_out.value]

    def dft(self, coeffs):
        return self._run("dft", coeffs)

    def idft(self, points):
        return self._run("idft", points)

    def coset_dft(self, coeffs):
        return self._run("coset_dft", coeffs)

    def coset_idft(self, points):
        return self._run("coset_idft", points)

    def h_coefficients(self, a, b, c):
        """groth16/src/prover.rs:36-47: the coefficients of H(X) from the R1CS evaluation vectors (three idft + three coset_dft,
        pointwise a*b - c, division by Z on the coset, coset_idft), all on the device in one call."""
        a, b, c = (_c(x).reshape(-1, 4) for x in (a, b, c))
        assert a.shape == b.shape == c.shape and a.shape[0] <= self.n
        out = np.zeros((self.n, 4), dtype=np.uint64)
        n_out = ctypes.c_size_t()
        _lib.check(_lib.lib().kgr_groth16_h(self.k, _u64(a), _u64(b), _u64(c), a.shape[0], _u64(out), ctypes.byref(n_out)))
        return out[: n_out.value]

    def transform_device(self, op, d_ptr):
        """In place on 2^k elements in device memory (raw pointer); for benchmarks."""
        _lib.check(_lib.lib().kgr_ntt_device(self.k, OPS[op], ctypes.c_void_p(d_ptr)))
