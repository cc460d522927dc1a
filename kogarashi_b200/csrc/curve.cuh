// curve.cuh — a = 0 short-Weierstrass group law for BN254 G1 and Grumpkin on the device.
//
// The reference accumulates buckets in homogeneous projective coordinates with a 9M+2S mixed
// add (zkstd/src/arithmetic/points/weierstrass.rs:63-97).  An MSM result is a group element with
// a unique normalised affine form, so any correct group law is bit-exact after to_affine; the
// device uses extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) whose
// mixed add is 8M+2S and needs no curve constant.  Every special case the reference branches on
// (identity operands :64-68, equal points -> doubling :76-78, opposite points -> identity
// :79-81) is handled explicitly because legal inputs hit them (identity CRS entries, repeated
// bases, P and -P in one bucket).
#pragma once
#include "field.cuh"

namespace kgr {

// bn254/src/g1.rs:18-22 / params.rs:8-12 : coordinates in Fq, scalars in Fr, b = 3
struct Bn254G1 {
    typedef FqP Base;
    typedef FrP Scalar;
    typedef Fp<FqP> Elem;
    static constexpr int ID = 0;
};
// grumpkin/src/curve.rs:11-15 / params.rs:4-19 : coordinates in Fr, scalars in Fq, b = -17
struct GrumpkinC {
    typedef FrP Base;
    typedef FqP Scalar;
    typedef Fp<FrP> Elem;
    static constexpr int ID = 1;
};
// bn254/src/g2.rs:15-21 / params.rs:14-56 : coordinates in Fq2 = Fq[u]/(u^2+1), scalars in Fr, b = 3/(9+u)
struct Bn254G2 {
    typedef FqP Base;
    typedef FrP Scalar;
    typedef Fp2<FqP> Elem;
    static constexpr int ID = 2;
};

// Device-side affine point, 64 bytes (128 for G2).  The reference carries a separate is_infinity flag
// (g1.rs:21); on the device the identity is encoded as (0, 0), which is on neither curve
// (b != 0), so the flag array is folded into the coordinates when bases are registered.
template <class C> struct alignas(16) AffinePt {
    typename C::Elem x, y;
};
template <class C> struct alignas(16) XyzzPt {
    typename C::Elem x, y, zz, zzz;
};

template <class C> KGR_HD bool affine_is_identity(const AffinePt<C> &p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
template <class C> KGR_HD bool xyzz_is_identity(const XyzzPt<C> &p) { return fp_is_zero(p.zz); }
template <class C> KGR_HD XyzzPt<C> xyzz_identity() {
    XyzzPt<C> r;
    r.x = El<typename C::Elem>::zero();
    r.y = El<typename C::Elem>::one();
    r.zz = El<typename C::Elem>::zero();
    r.zzz = El<typename C::Elem>::zero();
    return r;
}
template <class C> KGR_HD XyzzPt<C> xyzz_from_affine(const AffinePt<C> &p) {
    if (affine_is_identity(p)) return xyzz_identity<C>();
    XyzzPt<C> r;
    r.x = p.x;
    r.y = p.y;
    r.zz = El<typename C::Elem>::one();
    r.zzz = El<typename C::Elem>::one();
    return r;
}

// 2*(x, y) for an affine point (mdbl-2008-s-1, a = 0): 3M + 2S... (U^2, U*V, x*V, x^2, M^2, M*(..), W*y)
template <class C> KGR_HD XyzzPt<C> xyzz_dbl_affine(const AffinePt<C> &p) {
    typedef typename C::Elem E;
    if (affine_is_identity(p)) return xyzz_identity<C>();
    XyzzPt<C> r;
    E u = fp_dbl(p.y);
    E v = fp_sqr(u);
    E w = fp_mul(u, v);
    E s = fp_mul(p.x, v);
    E xx = fp_sqr(p.x);
    E m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2*P (dbl-2008-s-1, a = 0).  y == 0 would give zz = 0 = identity, which is the right answer
// for a 2-torsion point (neither curve has one: both groups have prime order).
template <class C> KGR_HD XyzzPt<C> xyzz_dbl(const XyzzPt<C> &p) {
    typedef typename C::Elem E;
    if (xyzz_is_identity(p)) return p;
    XyzzPt<C> r;
    E u = fp_dbl(p.y);
    E v = fp_sqr(u);
    E w = fp_mul(u, v);
    E s = fp_mul(p.x, v);
    E xx = fp_sqr(p.x);
    E m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = fp_mul(v, p.zz);
    r.zzz = fp_mul(w, p.zzz);
    return r;
}

// acc += q for an affine q (madd-2008-s): 8M + 2S on the generic path.
template <class C> KGR_HD void xyzz_madd(XyzzPt<C> &acc, const AffinePt<C> &q) {
    typedef typename C::Elem E;
    if (affine_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = xyzz_from_affine(q);
        return;
    }
    E u2 = fp_mul(q.x, acc.zz);
    E s2 = fp_mul(q.y, acc.zzz);
    E p = fp_sub(u2, acc.x);
    E r = fp_sub(s2, acc.y);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_dbl_affine(q);
        else acc = xyzz_identity<C>();
        return;
    }
    E pp = fp_sqr(p);
    E ppp = fp_mul(p, pp);
    E qq = fp_mul(acc.x, pp);
    E x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
    E y3 = fp_sub(fp_mul(r, fp_sub(qq, x3)), fp_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(acc.zz, pp);
    acc.zzz = fp_mul(acc.zzz, ppp);
}

// acc += q (add-2008-s): 12M + 2S on the generic path.
template <class C> KGR_HD void xyzz_add(XyzzPt<C> &acc, const XyzzPt<C> &q) {
    typedef typename C::Elem E;
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    E u1 = fp_mul(acc.x, q.zz);
    E u2 = fp_mul(q.x, acc.zz);
    E s1 = fp_mul(acc.y, q.zzz);
    E s2 = fp_mul(q.y, acc.zzz);
    E p = fp_sub(u2, u1);
    E r = fp_sub(s2, s1);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_dbl(acc);
        else acc = xyzz_identity<C>();
        return;
    }
    E pp = fp_sqr(p);
    E ppp = fp_mul(p, pp);
    E qq = fp_mul(u1, pp);
    E x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
    E y3 = fp_sub(fp_mul(r, fp_sub(qq, x3)), fp_mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(fp_mul(acc.zz, q.zz), pp);
    acc.zzz = fp_mul(fp_mul(acc.zzz, q.zzz), ppp);
}

#if defined(__CUDACC__)
// ---- quad-cooperative group law (device only) ----------------------------------------------------------------------------------------------
// The tail of the bucket reduction (deep fold levels, tree sums, the per-bit doubling chains) is a chain of dependent point operations on a
// handful of points: one lone warp needs ~6-8 us per general addition because its 14 field multiplications run one after the other.  Here the
// FOUR lanes of a quad hold the same operands, each computes one of up to four independent products of a round and the results are exchanged
// with shuffles: 4 rounds instead of 14 multiplications for an addition, 3 instead of 9 for a doubling.  Same formulas, same special cases and
// the same field elements as xyzz_add / xyzz_dbl (add-2008-s, dbl-2008-s-1), so the result is identical in every word.
// All four lanes of the quad must call with identical arguments; the result is returned in all four.
template <class E> __device__ __forceinline__ E quad_get(const E &v, unsigned mask, int src_lane) {
    E r;
#pragma unroll
    for (int w = 0; w < El<E>::WORDS; w++) El<E>::word(r, w) = __shfl_sync(mask, El<E>::word(v, w), src_lane);
    return r;
}
template <class E> __device__ __forceinline__ E quad_sel(int role, const E &a, const E &b, const E &c, const E &d) {
    E r;
#pragma unroll
    for (int w = 0; w < El<E>::WORDS; w++) {
        uint32_t lo = role & 1 ? El<E>::word(b, w) : El<E>::word(a, w), hi = role & 1 ? El<E>::word(d, w) : El<E>::word(c, w);
        El<E>::word(r, w) = role & 2 ? hi : lo;
    }
    return r;
}
template <class C> __device__ __forceinline__ XyzzPt<C> xyzz_dbl_quad(const XyzzPt<C> &p) {
    typedef typename C::Elem E;
    if (xyzz_is_identity(p)) return p;
    const int lane = (int)(threadIdx.x & 31u), role = lane & 3, qb = lane & 28;
    const unsigned mask = 0xFu << qb;
    XyzzPt<C> r;
    E u = fp_dbl(p.y);
    // round 1: v = u^2 | xx = x^2
    E m1 = fp_sqr(quad_sel(role, u, p.x, u, p.x));
    E v = quad_get(m1, mask, qb), xx = quad_get(m1, mask, qb + 1);
    E m = fp_add(fp_dbl(xx), xx);
    // round 2: w = u v | s = x v | zz3 = v zz | mm = m^2
    E m2 = fp_mul(quad_sel(role, u, p.x, v, m), quad_sel(role, v, v, p.zz, m));
    E w = quad_get(m2, mask, qb), sv = quad_get(m2, mask, qb + 1), mm = quad_get(m2, mask, qb + 3);
    r.zz = quad_get(m2, mask, qb + 2);
    r.x = fp_sub(mm, fp_dbl(sv));
    // round 3: m (s - x3) | w y | zzz3 = w zzz
    E m3 = fp_mul(quad_sel(role, m, w, w, w), quad_sel(role, fp_sub(sv, r.x), p.y, p.zzz, p.zzz));
    r.y = fp_sub(quad_get(m3, mask, qb), quad_get(m3, mask, qb + 1));
    r.zzz = quad_get(m3, mask, qb + 2);
    return r;
}
template <class C> __device__ __forceinline__ void xyzz_add_quad(XyzzPt<C> &acc, const XyzzPt<C> &q) {
    typedef typename C::Elem E;
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    const int lane = (int)(threadIdx.x & 31u), role = lane & 3, qb = lane & 28;
    const unsigned mask = 0xFu << qb;
    // round 1: u1 = x1 zz2 | u2 = x2 zz1 | s1 = y1 zzz2 | s2 = y2 zzz1
    E m1 = fp_mul(quad_sel(role, acc.x, q.x, acc.y, q.y), quad_sel(role, q.zz, acc.zz, q.zzz, acc.zzz));
    E u1 = quad_get(m1, mask, qb), u2 = quad_get(m1, mask, qb + 1), s1 = quad_get(m1, mask, qb + 2), s2 = quad_get(m1, mask, qb + 3);
    E p = fp_sub(u2, u1);
    E r = fp_sub(s2, s1);
    if (fp_is_zero(p)) {  // the same for all four lanes
        if (fp_is_zero(r)) acc = xyzz_dbl_quad(acc);
        else acc = xyzz_identity<C>();
        return;
    }
    // round 2: pp = p^2 | rr = r^2 | zz1 zz2 | zzz1 zzz2
    E m2 = fp_mul(quad_sel(role, p, r, acc.zz, acc.zzz), quad_sel(role, p, r, q.zz, q.zzz));
    E pp = quad_get(m2, mask, qb), rr = quad_get(m2, mask, qb + 1), zz12 = quad_get(m2, mask, qb + 2), zzz12 = quad_get(m2, mask, qb + 3);
    // round 3: ppp = p pp | qq = u1 pp | zz3 = zz12 pp
    E m3 = fp_mul(quad_sel(role, p, u1, zz12, zz12), pp);
    E ppp = quad_get(m3, mask, qb), qq = quad_get(m3, mask, qb + 1);
    acc.zz = quad_get(m3, mask, qb + 2);
    acc.x = fp_sub(fp_sub(rr, ppp), fp_dbl(qq));
    // round 4: r (qq - x3) | s1 ppp | zzz3 = zzz12 ppp
    E m4 = fp_mul(quad_sel(role, r, s1, zzz12, zzz12), quad_sel(role, fp_sub(qq, acc.x), ppp, ppp, ppp));
    acc.y = fp_sub(quad_get(m4, mask, qb), quad_get(m4, mask, qb + 1));
    acc.zzz = quad_get(m4, mask, qb + 2);
}
#endif

// XYZZ -> the reference's homogeneous projective (X : Y : Z), x = X/Z, y = Y/Z
// (zkstd/src/macros/curve/weierstrass/group.rs:106-110 identity = (0, 1, 0)).
template <class C> KGR_HD void xyzz_to_projective(const XyzzPt<C> &p, typename C::Elem out[3]) {
    typedef typename C::Elem E;
    if (xyzz_is_identity(p)) {
        out[0] = El<E>::zero();
        out[1] = El<E>::one();
        out[2] = El<E>::zero();
        return;
    }
    out[0] = fp_mul(p.x, p.zzz);
    out[1] = fp_mul(p.y, p.zz);
    out[2] = fp_mul(p.zz, p.zzz);
}

// XYZZ -> affine (x, y) or (0,0) for the identity: one inversion (macros/curve/weierstrass.rs:57-66 semantics)
template <class C> KGR_HD AffinePt<C> xyzz_to_affine(const XyzzPt<C> &p) {
    typedef typename C::Elem E;
    AffinePt<C> r;
    if (xyzz_is_identity(p)) {
        r.x = El<E>::zero();
        r.y = El<E>::zero();
        return r;
    }
    // 1/zzz ; 1/zz = zzz^-1 * ... : zz^3 = zzz^2  =>  1/zz = zz^2 / zzz^2 = (zz / zzz)^2 ... use one inversion of zz*zzz
    E t = fp_inv(fp_mul(p.zz, p.zzz));
    E izz = fp_mul(t, p.zzz);
    E izzz = fp_mul(t, p.zz);
    r.x = fp_mul(p.x, izz);
    r.y = fp_mul(p.y, izzz);
    return r;
}

}  // namespace kgr
