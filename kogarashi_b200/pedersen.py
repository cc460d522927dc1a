"""Mirror of nova's PedersenCommitment over the GPU MSM.

Reference (nova/src/pedersen.rs:5-21):
    pub struct PedersenCommitment<C: BNAffine> { g: Vec<C> }
    pub fn new<R: RngCore>(n: u64, rng: &mut R) -> Self        // 2^n + 1 random points (:10-13)
    pub(crate) fn commit(&self, m: &DenseVectors<C::Scalar>) -> C   // fold of sum + g_i * m_i, .into() affine (:15-20)
The reference's commit is a naive per-element scalar multiplication; the value it computes is
sum_i m_i * g_i over zip(m, g), which is exactly an MSM followed by to_affine.
"""
import ctypes
import os

import numpy as np

from . import _lib
from .msm import Bases, _c, _u64, SCALARS_MONTGOMERY


# bn254/src/fr.rs:11-16 (scalars of G1 / G2) and bn254/src/fq.rs:10-15 (scalars of Grumpkin)
_FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_SCALAR_MODULUS = {0: _FR, 1: _FQ, 2: _FR}


class PedersenCommitment:
    def __init__(self, curve, g, inf=None):
        """g: (k, 8) uint64 affine generators (x||y Montgomery) — uploaded once, like `ck` in nova."""
        self.curve = curve
        self.g = Bases(curve, g, inf) if not isinstance(g, Bases) else g

    @classmethod
    def new(cls, curve, n, rng=None):
        """2^n + 1 generators (`0..=1 << n`, pedersen.rs:11), g_i = G * Scalar::random(rng) (group.rs:39-41) with the k_i * G computed on
        the device (fixed-base window table).  `rng(nbytes) -> bytes` supplies the randomness (default os.urandom, the reference's OsRng);
        every k_i is 64 random bytes reduced modulo the scalar modulus, the distribution of `from_u512` (represent.rs:18-28,80-103)."""
        rng = rng or os.urandom
        count = (1 << n) + 1
        r = _SCALAR_MODULUS[curve]
        R = (1 << 256) % r
        k = np.zeros((count, 4), dtype=np.uint64)
        for i in range(count):
            v = (int.from_bytes(rng(64), "little") % r) * R % r          # Montgomery form, as the ABI takes scalars
            for j in range(4):
                k[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        from .msm import fixed_base_mul
        return cls(curve, fixed_base_mul(curve, k))

    @classmethod
    def new_benchmark(cls, curve, n, seed):
        """Synthetic key for benchmarks and tests ONLY: the k_i come from a public 64-bit splitmix64 stream on the device
        (kgr_bases_generate), so every discrete logarithm is computable and commitments under this key are NOT binding."""
        return cls(curve, Bases.generate(curve, (1 << n) + 1, seed))

    def commit(self, m, scalar_fmt=SCALARS_MONTGOMERY):
        """-> (9,) uint64: x[4], y[4], is_infinity."""
        m = _c(m).reshape(-1, 4)
        out = np.zeros(9, dtype=np.uint64)
        _lib.check(_lib.lib().kgr_pedersen_commit(self.g._h, _u64(m), scalar_fmt, m.shape[0], _u64(out)))
        return out
