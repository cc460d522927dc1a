"""Mirror of nova's PedersenCommitment over the GPU MSM.

Reference (nova/src/pedersen.rs:5-21):
    pub struct PedersenCommitment<C: BNAffine> { g: Vec<C> }
    pub fn new<R: RngCore>(n: u64, rng: &mut R) -> Self        // 2^n + 1 random points (:10-13)
    pub(crate) fn commit(&self, m: &DenseVectors<C::Scalar>) -> C   // fold of sum + g_i * m_i, .into() affine (:15-20)
The reference's commit is a naive per-element scalar multiplication; the value it computes is
sum_i m_i * g_i over zip(m, g), which is exactly an MSM followed by to_affine.
"""
import ctypes

import numpy as np

from . import _lib
from .msm import Bases, _c, _u64, SCALARS_MONTGOMERY


class PedersenCommitment:
    def __init__(self, curve, g, inf=None):
        """g: (k, 8) uint64 affine generators (x||y Montgomery) — uploaded once, like `ck` in nova."""
        self.curve = curve
        self.g = Bases(curve, g, inf) if not isinstance(g, Bases) else g

    @classmethod
    def new(cls, curve, n, seed=1):
        """2^n + 1 generators (`0..=1 << n`, pedersen.rs:11), each k_i * G with k_i drawn on the device."""
        return cls(curve, Bases.generate(curve, (1 << n) + 1, seed))

    def commit(self, m, scalar_fmt=SCALARS_MONTGOMERY):
        """-> (9,) uint64: x[4], y[4], is_infinity."""
        m = _c(m).reshape(-1, 4)
        out = np.zeros(9, dtype=np.uint64)
        _lib.check(_lib.lib().kgr_pedersen_commit(self.g._h, _u64(m), scalar_fmt, m.shape[0], _u64(out)))
        return out
