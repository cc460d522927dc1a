"""oracle/pyref.py — TEST INFRASTRUCTURE ("Oracle B"), NOT PRODUCT CODE.

Independent textbook big-integer model of the same maths the reference implements with limbs:
affine chord-and-tangent group law, `pow(x, -1, p)` inversion, plain integer scalars.  It shares
no structure with oracle/zkstd_oracle.hpp (no Montgomery form, no projective coordinates, no
windows), so agreement of the two on the golden vectors is meaningful evidence for both.

Reference anchors (paths relative to /root/reference):
  moduli            bn254/src/fq.rs:10-15, bn254/src/fr.rs:11-16
  G1                bn254/src/params.rs:8-12        y^2 = x^3 + 3, generator (1, 2)
  Grumpkin          grumpkin/src/params.rs:4-19     y^2 = x^3 - 17 over Fr, generator (1, sqrt(-16))
  MSM semantics     groth16/src/msm.rs:6-48         sum over the first min(len) pairs
  sampler           zkstd/src/arithmetic/limbs/bits_256/represent.rs:18-28,80-103
"""

FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R256 = 1 << 256

BN254_G1 = 0
GRUMPKIN = 1


class CurveModel:
    def __init__(self, name, p, r, b, gx, gy):
        self.name, self.p, self.r, self.b, self.g = name, p, r, b % p, (gx % p, gy % p)
        assert self.on_curve(self.g)

    def on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return (y * y - (x * x * x + self.b)) % self.p == 0

    def neg(self, pt):
        return None if pt is None else (pt[0], (-pt[1]) % self.p)

    def add(self, a, b):
        p = self.p
        if a is None:
            return b
        if b is None:
            return a
        if a[0] == b[0]:
            if (a[1] + b[1]) % p == 0:
                return None
            lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, p) % p
        else:
            lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, p) % p
        x = (lam * lam - a[0] - b[0]) % p
        return (x, (lam * (a[0] - x) - a[1]) % p)

    def mul(self, pt, k):
        k %= self.r
        acc = None
        while k:
            if k & 1:
                acc = self.add(acc, pt)
            pt = self.add(pt, pt)
            k >>= 1
        return acc

    def msm(self, points, scalars):
        """points: list of (x, y) or None; scalars: ints.  zip semantics like msm.rs:25."""
        acc = None
        for pt, k in zip(points, scalars):
            acc = self.add(acc, self.mul(pt, k))
        return acc


def _grumpkin_gy():
    # grumpkin/src/params.rs:6-11 stores GENERATOR_Y in Montgomery form (R = 2^256 mod r)
    limbs = [0x11B2DFF1448C41D8, 0x23D3446F21C77DC3, 0xAA7B8CF435DFAFBB, 0x14B34CF69DC25D68]
    v = sum(l << (64 * i) for i, l in enumerate(limbs))
    return v * pow(R256, -1, FR) % FR


CURVES = {
    BN254_G1: CurveModel("bn254_g1", FQ, FR, 3, 1, 2),
    GRUMPKIN: CurveModel("grumpkin", FR, FQ, -17, 1, _grumpkin_gy()),
}


# ---- limb <-> int helpers (Montgomery R = 2^256) --------------------------------------------
def limbs_to_int(limbs):
    return sum(int(l) << (64 * i) for i, l in enumerate(limbs))


def int_to_limbs(v, n=4):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


_RINV = {}


def to_mont(v, p):
    return v * R256 % p


def from_mont(v, p):
    if p not in _RINV:
        _RINV[p] = pow(R256, -1, p)
    return v * _RINV[p] % p


def from_u512(words, p):
    """represent.rs:18-28: (lo*R^2 + hi*R^3) * R^-1 in Montgomery form == (lo + 2^256 hi) mod p."""
    lo = limbs_to_int(words[:4])
    hi = limbs_to_int(words[4:])
    return (lo + (hi << 256)) % p


class XorShift128:
    """rand_xorshift's XorShiftRng restated from its public algorithm (see zkstd_oracle.hpp)."""

    def __init__(self, seed16: bytes):
        s = [int.from_bytes(seed16[4 * i:4 * i + 4], "little") for i in range(4)]
        if not any(s):
            s = [0x0BAD5EED] * 4
        self.x, self.y, self.z, self.w = s

    def next_u32(self):
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return (hi << 32) | lo

    def random_field(self, p):
        return from_u512([self.next_u64() for _ in range(8)], p)
