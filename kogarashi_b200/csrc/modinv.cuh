// modinv.cuh — modular inversion by batched divsteps ("safegcd", Bernstein & Yang 2019), for Fq / Fr.
//
// The reference inverts with Fermat's little theorem, a^(p-2) by square-and-multiply over all 256 exponent bits
// (zkstd/src/arithmetic/limbs/bits_256/normal.rs:256-287): ~380 field multiplications.  The inverse is unique, so any
// algorithm returns the same bits; this one costs about 25 multiplications' worth of time and makes batched affine
// additions affordable with small batches.
//
// divstep:   if delta > 0 and g odd:  (delta, f, g) <- (1 - delta, g, (g - f) / 2)
//            else:                    (delta, f, g) <- (1 + delta, f, (g + (g odd ? f : 0)) / 2)
// starting from (1, p, a) it ends with g = 0, f = +-1.  Thirty divsteps depend only on the low 30 bits of f and g and
// are summarised by a 2x2 integer matrix T with  2^30 (f', g') = T (f, g);  the same matrix updates (d, e) with
// d a = f, e a = g (mod p), dividing by 2^30 modulo p.  Numbers are nine signed 30-bit limbs.  Plain C++ (the same
// code runs on the host and on the device); the loop runs until g = 0.
#pragma once
#include "field.cuh"

namespace kgr {

struct Limbs30 {
    int32_t v[9];
};

template <class P> KGR_HD Limbs30 modulus30() {
    Limbs30 m;
    uint64_t acc = 0;
    int bits = 0, k = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc |= (uint64_t)P::mod(i) << bits;
        bits += 32;
        while (bits >= 30 && k < 8) {
            m.v[k++] = (int32_t)(acc & 0x3fffffffu);
            acc >>= 30;
            bits -= 30;
        }
    }
    m.v[8] = (int32_t)acc;
    return m;
}
KGR_HD Limbs30 to_limbs30(const uint32_t a[8]) {
    Limbs30 r;
    uint64_t acc = 0;
    int bits = 0, k = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc |= (uint64_t)a[i] << bits;
        bits += 32;
        while (bits >= 30 && k < 8) {
            r.v[k++] = (int32_t)(acc & 0x3fffffffu);
            acc >>= 30;
            bits -= 30;
        }
    }
    r.v[8] = (int32_t)acc;
    return r;
}
// value in [0, 2^256) with non-negative limbs -> 8 x 32-bit
KGR_HD void from_limbs30(const Limbs30 &a, uint32_t out[8]) {
    uint64_t acc = 0;
    int bits = 0, k = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        acc |= (uint64_t)(uint32_t)a.v[i] << bits;
        bits += 30;
        while (bits >= 32 && k < 8) {
            out[k++] = (uint32_t)acc;
            acc >>= 32;
            bits -= 32;
        }
    }
    if (k < 8) out[k] = (uint32_t)acc;
}

// -p^-1 is not needed here; this is p^-1 mod 2^30 by Newton iteration on the low limb (p odd)
KGR_HD uint32_t inv30(uint32_t p0) {
    uint32_t x = p0;  // correct to 3 bits
    x *= 2 - p0 * x;
    x *= 2 - p0 * x;
    x *= 2 - p0 * x;
    x *= 2 - p0 * x;
    return x & 0x3fffffffu;
}

struct Trans {
    int32_t u, v, q, r;
};

// 30 divsteps on the low bits.  f0, g0: low 32 bits of f and g (f odd).  Returns the new delta.
KGR_HD int32_t divsteps30(int32_t delta, uint32_t f0, uint32_t g0, Trans &t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 1
    for (int i = 0; i < 30; i++) {
        uint32_t odd = 0u - (g & 1u);                                   // all ones if g is odd
        uint32_t swap = odd & (0u - (uint32_t)(delta > 0));            // delta > 0 and g odd
        // swap: (f, g) <- (g, -f), rows (u v | q r) <- (q r | -u -v), delta <- -delta
        uint32_t nf = (f & ~swap) | (g & swap), ng = (g & ~swap) | ((0u - f) & swap);
        uint32_t nu = (u & ~swap) | (q & swap), nq = (q & ~swap) | ((0u - u) & swap);
        uint32_t nv = (v & ~swap) | (r & swap), nr = (r & ~swap) | ((0u - v) & swap);
        delta = swap ? -delta : delta;
        f = nf; g = ng; u = nu; q = nq; v = nv; r = nr;
        delta += 1;
        g += f & odd;
        q += u & odd;
        r += v & odd;
        g >>= 1;                                                          // exact: g is even now (low bits only)
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return delta;
}

// (f, g) <- T (f, g) / 2^30   (exact division)
KGR_HD void update_fg30(Limbs30 &f, Limbs30 &g, const Trans &t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    int64_t cf = u * f.v[0] + v * g.v[0];
    int64_t cg = q * f.v[0] + r * g.v[0];
    cf >>= 30;  // low 30 bits are zero by construction
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        int64_t fi = f.v[i], gi = g.v[i];
        cf += u * fi + v * gi;
        cg += q * fi + r * gi;
        f.v[i - 1] = (int32_t)(cf & 0x3fffffff);
        g.v[i - 1] = (int32_t)(cg & 0x3fffffff);
        cf >>= 30;
        cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}

// (d, e) <- T (d, e) / 2^30 mod p, keeping d, e in (-2p, p)
KGR_HD void update_de30(Limbs30 &d, Limbs30 &e, const Trans &t, const Limbs30 &m, uint32_t m_inv30) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;  // -1 if negative
    // adding p to a negative d / e first: multiples of p contributed through the matrix
    int32_t md = (t.u & sd) + (t.v & se);
    int32_t me = (t.q & sd) + (t.r & se);
    int64_t cd = u * d.v[0] + v * e.v[0];
    int64_t ce = q * d.v[0] + r * e.v[0];
    // choose the multiples of p that clear the low 30 bits; correction term in (-2^30, 0]
    md -= (int32_t)((m_inv30 * (uint32_t)cd + (uint32_t)md) & 0x3fffffffu);
    me -= (int32_t)((m_inv30 * (uint32_t)ce + (uint32_t)me) & 0x3fffffffu);
    cd += (int64_t)m.v[0] * md;
    ce += (int64_t)m.v[0] * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        int64_t di = d.v[i], ei = e.v[i];
        cd += u * di + v * ei + (int64_t)m.v[i] * md;
        ce += q * di + r * ei + (int64_t)m.v[i] * me;
        d.v[i - 1] = (int32_t)(cd & 0x3fffffff);
        e.v[i - 1] = (int32_t)(ce & 0x3fffffff);
        cd >>= 30;
        ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}

// a^-1 mod p for a canonical integer 0 < a < p (limbs 8 x 32); 0 maps to 0.
template <class P> KGR_HD void modinv_int(const uint32_t a[8], uint32_t out[8]) {
    const Limbs30 m = modulus30<P>();
    const uint32_t m_inv30 = inv30((uint32_t)m.v[0]);
    Limbs30 f = m, g = to_limbs30(a), d, e;
#pragma unroll
    for (int i = 0; i < 9; i++) d.v[i] = e.v[i] = 0;
    e.v[0] = 1;
    int32_t delta = 1;
#pragma unroll 1
    for (int batch = 0; batch < 40; batch++) {
        int32_t gz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) gz |= g.v[i];
        if (gz == 0) break;
        Trans t;
        // low 32 bits of f and g from the two lowest limbs
        uint32_t f0 = (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), g0 = (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30);
        delta = divsteps30(delta, f0, g0, t);
        update_de30(d, e, t, m, m_inv30);
        update_fg30(f, g, t);
    }
    // f = +-1 (or +-gcd); inverse = sign(f) * d, brought into [0, p)
    const int32_t sf = f.v[8] >> 31;  // -1 if f negative
    // d in (-2p, p): add p while negative (at most twice), then negate if f < 0
    for (int rep = 0; rep < 2; rep++) {
        int32_t neg = d.v[8] >> 31;
        int64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (int64_t)d.v[i] + (m.v[i] & neg);
            d.v[i] = (int32_t)(c & 0x3fffffff);
            c >>= 30;
        }
        d.v[8] = (int32_t)(c + d.v[8] + (m.v[8] & neg));  // the top limb stays signed
    }
    if (sf) {  // d <- p - d  (d in [0, p)); 0 stays 0
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) nz |= d.v[i];
        if (nz) {
            int64_t c = 0;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                c += (int64_t)m.v[i] - d.v[i];
                d.v[i] = (int32_t)(c & 0x3fffffff);
                c >>= 30;
            }
        }
    }
    from_limbs30(d, out);
}

// Montgomery-form inverse: a R -> a^-1 R.  Zero maps to zero (callers test for zero where the reference returns None).
template <class P> KGR_HD Fp<P> fp_inv_fast(const Fp<P> &a) {
    Fp<P> x, r3;
    modinv_int<P>(a.v, x.v);  // (a R)^-1 = a^-1 R^-1
#pragma unroll
    for (int i = 0; i < 8; i++) r3.v[i] = P::r3(i);
    return fp_mul(x, r3);     // a^-1 R^-1 * R^3 * R^-1 = a^-1 R
}

template <class P> KGR_HD Fp2<P> fp_inv_fast(const Fp2<P> &a) {
    Fp<P> t = fp_inv_fast(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
    return Fp2<P>{fp_mul(t, a.c0), fp_mul(t, fp_neg(a.c1))};
}

}  // namespace kgr
