"""oracle/groth16_ref.py — TEST INFRASTRUCTURE: big-integer restatement of the reference's Groth16
setup and prover for its example circuit (BASELINE config #4), so that the GPU-routed G1 commitments
can be compared byte for byte with the all-CPU reference computation under fixed randomness.

PARITY PIN STATUS: the Rust reference cannot be executed here and records no proof bytes anywhere, so
agreement of these bytes with the Rust binary is UNPINNED.  What pins this file: the proof it produces
satisfies the Groth16 verification equation (checked in the exponent with the known toxic waste,
tests/test_groth16.py) and every point is checked against the independent exponent formula.

Follows (paths relative to /root/reference):
  groth16/examples/simple.rs:32-50      DummyCircuit: x^3 + x + 5 == o
  zkstd/src/circuit/gadget/field.rs:14-76,183-185   instance / constant / mul / add / enforce_eq
  zkstd/src/r1cs.rs:29-167              m, l, m_l_1, evaluate, z_vectors, gates
  groth16/src/fft.rs:27-160             domain, (coset) dft / idft, z, divide_by_z_on_coset
  groth16/src/poly.rs:61-63             Coefficients::new strips trailing zeros
  groth16/src/zksnark.rs:17-200         setup, eval, eval_at_tau
  groth16/src/prover.rs:20-99           create_proof (order of RNG draws, assembly of A, B, C)
  bn254/src/params.rs:14-57             G2 generator and b; bn254/src/fr.rs:18,53-66 S, ROOT_OF_UNITY, generator 7
Scalars are plain integers mod r; curve points are affine tuples or None (identity); the DFTs are
evaluated from their definition (the field arithmetic is exact, so the recursive radix-2 schedule of
fft.rs:166-218 yields the same values).
"""
from . import pyref as B

R = B.FR
Q = B.FQ
S = 28
ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C  # fr.rs:58-63
MULT_GEN = 7  # fr.rs:18


# ---- Fq2 = Fq[u] / (u^2 + 1) and G2 -----------------------------------------------------------------
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


G2_GEN = ((0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED, 0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
          (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA, 0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B))
G2_B = (0x2B149D40CEB8AAAE81BE18991BE06AC3B5B4C5E559DBEFA33267E6DC24A138E5, 0x009713B03AF0FED4CD2CAFADEED8FDF4A74FA084E52D1852E4A2BD0685C315D2)


def g2_on_curve(p):
    if p is None:
        return True
    x, y = p
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), G2_B)


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if f2_add(a[1], b[1]) == (0, 0):
            return None
        x2 = f2_mul(a[0], a[0])
        lam = f2_mul(f2_add(f2_add(x2, x2), x2), f2_inv(f2_add(a[1], a[1])))
    else:
        lam = f2_mul(f2_sub(b[1], a[1]), f2_inv(f2_sub(b[0], a[0])))
    x = f2_sub(f2_sub(f2_mul(lam, lam), a[0]), b[0])
    return (x, f2_sub(f2_mul(lam, f2_sub(a[0], x)), a[1]))


def g2_mul(p, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


G1 = B.CURVES[B.BN254_G1]
assert g2_on_curve(G2_GEN)


# ---- R1CS of the example (field.rs gadgets restated on sparse rows {wire: coeff}) --------------------
class R1cs:
    """Wires: ('x', i) instance, ('w', i) witness; ('x', 0) is the constant one (r1cs.rs:177-189)."""

    def __init__(self):
        self.a, self.b, self.c, self.x, self.w = [], [], [], [1], []

    def value(self, row):
        return sum(co * (self.x[i] if kind == "x" else self.w[i]) for (kind, i), co in row.items()) % R

    def instance(self, v):
        self.x.append(v % R)
        return {("x", len(self.x) - 1): 1}

    def witness(self, v):
        self.w.append(v % R)
        return {("w", len(self.w) - 1): 1}

    @staticmethod
    def constant(c):
        return {("x", 0): c % R}

    @staticmethod
    def row_add(p, q):
        out = dict(p)
        for k, v in q.items():
            out[k] = (out.get(k, 0) + v) % R
        return out

    def mul(self, p, q):  # field.rs:48-62 (no constant operands in the example)
        z = self.witness(self.value(p) * self.value(q))
        self.a.append(p), self.b.append(q), self.c.append(z)
        return z

    def add(self, p, q):  # field.rs:64-76 -> r1cs.rs:117-124: (x + y) * 1 = z
        z = self.witness(self.value(p) + self.value(q))
        self.a.append(self.row_add(p, q)), self.b.append(self.constant(1)), self.c.append(z)
        return z

    def enforce_eq(self, p, q):  # field.rs:183-185: x * 1 = y
        self.a.append(p), self.b.append(self.constant(1)), self.c.append(q)

    m = property(lambda self: len(self.a))
    l = property(lambda self: len(self.x))
    m_l_1 = property(lambda self: len(self.w))

    def evaluate(self):
        return ([self.value(r) for r in self.a], [self.value(r) for r in self.b], [self.value(r) for r in self.c])

    def columns(self, mat):
        """matrix.rs:15-29 x_and_w: per wire, the list of (coeff, constraint index)."""
        xs, ws = [[] for _ in range(self.l)], [[] for _ in range(self.m_l_1)]
        for i, row in enumerate(mat):
            for (kind, k), co in row.items():
                (xs if kind == "x" else ws)[k].append((co, i))
        return xs, ws


def example_circuit(x, o):
    """groth16/examples/simple.rs:32-50."""
    cs = R1cs()
    xv, ov = cs.instance(x), cs.instance(o)
    c5 = cs.constant(5)
    sym1 = cs.mul(xv, xv)
    y = cs.mul(sym1, xv)
    sym2 = cs.add(y, xv)
    cs.enforce_eq(cs.row_add(sym2, c5), ov)
    return cs


def chain_circuit(steps, x0):
    """The example's function iterated (the map nova/src/test.rs:16-29 iterates): x_{i+1} = x_i^3 + x_i + 5,
    three constraints per step, public input x0 and public output x_steps (enforced like simple.rs:47)."""
    out = x0 % R
    for _ in range(steps):
        out = (out * out * out + out + 5) % R
    cs = R1cs()
    xv, ov = cs.instance(x0), cs.instance(out)
    c5 = cs.constant(5)
    cur = xv
    for _ in range(steps):
        sym1 = cs.mul(cur, cur)
        y = cs.mul(sym1, cur)
        sym2 = cs.add(y, cur)
        cur = cs.row_add(sym2, c5)
    cs.enforce_eq(cur, ov)
    return cs, out


def lagrange_at(n, omega, tau):
    """L_j(tau) for the size-n domain in closed form, = idft of the powers of tau (zksnark.rs:60): omega^j (tau^n - 1) / (n (tau - omega^j))."""
    zt = (pow(tau, n, R) - 1) % R
    ninv = pow(n, -1, R)
    w, ws, dens = 1, [], []
    for _ in range(n):
        ws.append(w)
        dens.append((tau - w) % R)
        w = w * omega % R
    pref, acc = [], 1
    for d in dens:  # batch inversion
        pref.append(acc)
        acc = acc * d % R
    inv = pow(acc, -1, R)
    out = [0] * n
    for j in range(n - 1, -1, -1):
        out[j] = ws[j] * zt % R * ninv % R * (inv * pref[j] % R) % R
        inv = inv * dens[j] % R
    return out


def crs_exponents(cs, rng):
    """Discrete logs of every G1 CRS element of zksnark.rs:17-127 for constraint system `cs` (so that the points can be produced by
    any fixed-base multiplication), plus the toxic waste and the per-variable (u, v, w)(tau)."""
    k = max(1, (1 << (cs.m - 1).bit_length()).bit_length() - 1)
    n = 1 << k
    omega = pow(ROOT_OF_UNITY, 1 << (S - k), R)
    alpha, beta, gamma, delta, tau = (rng.random_field(R) for _ in range(5))
    gamma_inv, delta_inv = pow(gamma, -1, R), pow(delta, -1, R)
    coeff = (pow(tau, n, R) - 1) * delta_inv % R
    h, p = [], 1
    for _ in range(cs.m - 1):
        h.append(p * coeff % R)
        p = p * tau % R
    lagr = lagrange_at(n, omega, tau)
    (ax, aw), (bx, bw), (cx, cw) = cs.columns(cs.a), cs.columns(cs.b), cs.columns(cs.c)
    ev = lambda col: sum(lagr[idx] * co for co, idx in col) % R
    a, b, ic, l, uvw = [], [], [], [], []
    for cols, ext, inv in (((ax, bx, cx), ic, gamma_inv), ((aw, bw, cw), l, delta_inv)):
        for ca, cb, cc in zip(*cols):
            at, bt, ct = ev(ca), ev(cb), ev(cc)
            uvw.append((at, bt, ct))
            a.append(at)
            b.append(bt)
            ext.append((at * beta + bt * alpha + ct) * inv % R)
    trap = dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, tau=tau)
    return dict(k=k, n=n, h=h, a=a, b_g1=b, ic=ic, l=l), trap, uvw


def expected_exponents(trap, uvw, inputs, aux, q, n, r, s):
    """Discrete logs of proof.a, proof.b and proof.c (prover.rs:75-92) and the Groth16 equation in the exponent."""
    al, be, ga, de, tau = (trap[k] for k in ("alpha", "beta", "gamma", "delta", "tau"))
    z = list(inputs) + list(aux)
    l = len(inputs)
    u = sum(zi * t[0] for zi, t in zip(z, uvw)) % R
    v = sum(zi * t[1] for zi, t in zip(z, uvw)) % R
    ht, p = 0, 1
    for qi in q:
        ht = (ht + qi * p) % R
        p = p * tau % R
    ht = ht * (pow(tau, n, R) - 1) % R
    a_exp = (al + r * de + u) % R
    b_exp = (be + s * de + v) % R
    aux_part = sum(zi * (be * t[0] + al * t[1] + t[2]) for zi, t in zip(z[l:], uvw[l:])) % R
    in_part = sum(zi * (be * t[0] + al * t[1] + t[2]) for zi, t in zip(z[:l], uvw[:l])) % R
    c_exp = ((aux_part + ht) * pow(de, -1, R) + s * a_exp + r * b_exp - r * s * de) % R
    pairing_ok = (a_exp * b_exp - (al * be + in_part + c_exp * de)) % R == 0
    return a_exp, b_exp, c_exp, pairing_ok


# ---- radix-2 domain (fft.rs) -----------------------------------------------------------------------
class Fft:
    def __init__(self, k):
        self.n = 1 << k
        self.omega = pow(ROOT_OF_UNITY, 1 << (S - k), R)
        assert pow(self.omega, self.n, R) == 1 and pow(self.omega, self.n // 2, R) != 1

    def dft(self, coeffs):
        c = list(coeffs) + [0] * (self.n - len(coeffs))
        return [sum(cj * pow(self.omega, i * j, R) for j, cj in enumerate(c)) % R for i in range(self.n)]

    def idft(self, points):
        p = list(points) + [0] * (self.n - len(points))
        wi, ninv = pow(self.omega, -1, R), pow(self.n, -1, R)
        return strip([sum(pj * pow(wi, i * j, R) for j, pj in enumerate(p)) * ninv % R for i in range(self.n)])

    def coset_dft(self, coeffs):
        return self.dft([c * pow(MULT_GEN, j, R) % R for j, c in enumerate(coeffs)])

    def coset_idft(self, points):
        gi = pow(MULT_GEN, -1, R)
        return strip([c * pow(gi, j, R) % R for j, c in enumerate(self.idft(points))])

    def z(self, tau):
        return (pow(tau, self.n, R) - 1) % R

    def z_on_coset(self):
        return (pow(MULT_GEN, self.n, R) - 1) % R


def strip(c):
    """poly.rs:61-63,133-138: Coefficients::new drops trailing zeros."""
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


# ---- setup and prover ------------------------------------------------------------------------------
def setup(rng):
    """zksnark.rs:17-127.  rng: oracle.pyref.XorShift128 (Fr::random = from_u512 of 8 next_u64)."""
    cs = example_circuit(0, 0)  # C::default()
    k = (1 << (cs.m - 1).bit_length()).bit_length() - 1
    fft = Fft(k)
    alpha, beta, gamma, delta, tau = (rng.random_field(R) for _ in range(5))
    gamma_inv, delta_inv = pow(gamma, -1, R), pow(delta, -1, R)
    powers = [pow(tau, i, R) for i in range(cs.m)]
    coeff = fft.z(tau) * delta_inv % R
    h = [G1.mul(G1.g, p * coeff % R) for p in powers[: cs.m - 1]]
    lagr = fft.idft(powers)
    lagr = lagr + [0] * (fft.n - len(lagr))
    n_var = cs.l + cs.m_l_1
    P = dict(a=[None] * n_var, b_g1=[None] * n_var, b_g2=[None] * n_var, ic=[None] * cs.l, l=[None] * cs.m_l_1, h=h)
    trap = dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, tau=tau)
    uvw = []  # (u_i(tau), v_i(tau), w_i(tau)) per variable, for the exponent checks

    def eval_at_tau(col):
        return sum(lagr[idx] * co for co, idx in col) % R

    (ax, aw), (bx, bw), (cx, cw) = cs.columns(cs.a), cs.columns(cs.b), cs.columns(cs.c)
    for off, cols, ext, inv in ((0, (ax, bx, cx), P["ic"], gamma_inv), (cs.l, (aw, bw, cw), P["l"], delta_inv)):
        for i, (ca, cb, cc) in enumerate(zip(*cols)):
            at, bt, ct = eval_at_tau(ca), eval_at_tau(cb), eval_at_tau(cc)
            uvw.append((at, bt, ct))
            if at:
                P["a"][off + i] = G1.mul(G1.g, at)
            if bt:
                P["b_g1"][off + i] = G1.mul(G1.g, bt)
                P["b_g2"][off + i] = g2_mul(G2_GEN, bt)
            ext[i] = G1.mul(G1.g, (at * beta + bt * alpha + ct) * inv % R)
    P["vk"] = dict(alpha_g1=G1.mul(G1.g, alpha), beta_g1=G1.mul(G1.g, beta), beta_g2=g2_mul(G2_GEN, beta), gamma_g2=g2_mul(G2_GEN, gamma),
                   delta_g1=G1.mul(G1.g, delta), delta_g2=g2_mul(G2_GEN, delta), ic=P["ic"])
    return P, trap, uvw


def g1_msm(points, scalars):
    return G1.msm(points, scalars)  # zip semantics like msm.rs:25


def g2_msm(points, scalars):
    acc = None
    for p, k in zip(points, scalars):
        acc = g2_add(acc, g2_mul(p, k)) if p is not None else acc
    return acc


def witness_scalars(x, o):
    """Everything create_proof feeds to the MSMs (prover.rs:25-56): q coefficients, input and aux assignments."""
    cs = example_circuit(x, o)
    k = (1 << (cs.m - 1).bit_length()).bit_length() - 1
    fft = Fft(k)
    a, b, c = cs.evaluate()
    a, b, c = (fft.coset_dft(fft.idft(v)) for v in (a, b, c))
    hq = [(ai * bi - ci) % R for ai, bi, ci in zip(a, b, c)]
    zi = pow(fft.z_on_coset(), -1, R)
    q = fft.coset_idft([v * zi % R for v in hq])
    return dict(q=q, inputs=list(cs.x), aux=list(cs.w), l=cs.l)


def create_proof(P, rng, x, o):
    """prover.rs:20-99 with every MSM on the CPU.  Returns (proof, internals)."""
    ws = witness_scalars(x, o)
    q, inputs, aux, l = ws["q"], ws["inputs"], ws["aux"], ws["l"]
    vk = P["vk"]
    q_pt = g1_msm(P["h"], q)
    l_pt = g1_msm(P["l"], aux)
    a_answer = G1.add(g1_msm(P["a"], inputs), g1_msm(P["a"][l:], aux))
    b1_answer = G1.add(g1_msm(P["b_g1"], inputs), g1_msm(P["b_g1"][l:], aux))
    b2_answer = g2_add(g2_msm(P["b_g2"], inputs), g2_msm(P["b_g2"][l:], aux))
    r, s = rng.random_field(R), rng.random_field(R)
    g_a = G1.add(G1.add(G1.mul(vk["delta_g1"], r), vk["alpha_g1"]), a_answer)
    g_b = g2_add(g2_add(g2_mul(vk["delta_g2"], s), vk["beta_g2"]), b2_answer)
    g_c = G1.add(G1.add(G1.mul(G1.mul(vk["delta_g1"], r), s), G1.mul(vk["alpha_g1"], s)), G1.mul(vk["beta_g1"], r))
    g_c = G1.add(g_c, G1.mul(a_answer, s))
    g_c = G1.add(g_c, G1.mul(b1_answer, r))
    g_c = G1.add(g_c, G1.add(q_pt, l_pt))
    return dict(a=g_a, b=g_b, c=g_c), dict(r=r, s=s, **ws)


# ---- encodings --------------------------------------------------------------------------------------
def fq_mont_bytes(v):
    return (v * B.R256 % Q).to_bytes(32, "little")


def encode_g1(p):
    """SCALE derive on G1Affine {x: Fq, y: Fq, is_infinity: bool} (bn254/src/g1.rs:17-22): Montgomery limbs LE + flag;
    identity = (0, R, true) (macros/curve/weierstrass/group.rs:22-26)."""
    if p is None:
        return fq_mont_bytes(0) + fq_mont_bytes(1) + b"\x01"
    return fq_mont_bytes(p[0]) + fq_mont_bytes(p[1]) + b"\x00"


def encode_g2(p):
    if p is None:
        return fq_mont_bytes(0) * 2 + fq_mont_bytes(1) + fq_mont_bytes(0) + b"\x01"
    return fq_mont_bytes(p[0][0]) + fq_mont_bytes(p[0][1]) + fq_mont_bytes(p[1][0]) + fq_mont_bytes(p[1][1]) + b"\x00"


def proof_bytes(proof):
    """a.encode() || b.encode() || c.encode()  (65 + 129 + 65 bytes; groth16/src/proof.rs:7-11 has no Encode of its own)."""
    return encode_g1(proof["a"]) + encode_g2(proof["b"]) + encode_g1(proof["c"])


def exponent_check(trap, uvw, internals, proof):
    """The Groth16 equation e(A,B) = e(alpha,beta) e(sum x_i IC_i, gamma) e(C, delta) and the three points, checked in
    the exponent with the toxic waste (no pairing needed)."""
    al, be, ga, de, tau = (trap[k] for k in ("alpha", "beta", "gamma", "delta", "tau"))
    z = internals["inputs"] + internals["aux"]
    l = internals["l"]
    r, s = internals["r"], internals["s"]
    u = sum(zi * t[0] for zi, t in zip(z, uvw)) % R
    v = sum(zi * t[1] for zi, t in zip(z, uvw)) % R
    ht = sum(qi * pow(tau, i, R) for i, qi in enumerate(internals["q"])) * (pow(tau, 4, R) - 1) % R
    a_exp = (al + r * de + u) % R
    b_exp = (be + s * de + v) % R
    aux_part = sum(zi * (be * t[0] + al * t[1] + t[2]) for zi, t in zip(z[l:], uvw[l:])) % R
    in_part = sum(zi * (be * t[0] + al * t[1] + t[2]) for zi, t in zip(z[:l], uvw[:l])) % R
    c_exp = ((aux_part + ht) * pow(de, -1, R) + s * a_exp + r * b_exp - r * s * de) % R
    ok_points = proof["a"] == G1.mul(G1.g, a_exp) and proof["b"] == g2_mul(G2_GEN, b_exp) and proof["c"] == G1.mul(G1.g, c_exp)
    ok_pairing = (a_exp * b_exp - (al * be + in_part + c_exp * de)) % R == 0
    return ok_points, ok_pairing
