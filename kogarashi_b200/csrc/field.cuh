// field.cuh — 254-bit prime-field arithmetic on 8 x 32-bit limbs for sm_100a.
//
// Values are Montgomery residues a*2^256 mod p kept fully reduced in [0, p), i.e. the same
// numbers the reference stores in its [u64; 4] limbs (zkstd/src/arithmetic/limbs/bits_256/
// normal.rs); a 4 x u64 limb array reinterpreted as 8 x u32 (little endian) is bit-identical.
//
// fp_mul is an interleaved (CIOS-style) Montgomery product on even/odd-aligned accumulators:
// every 32x32->64 partial product is one mad.lo.cc/madc.hi.cc pair that ptxas fuses into a
// single IMAD.WIDE.U32(.X) with a predicate carry, so a product costs 128 wide IMADs + 8 IMAD
// (the reference's mul+mont, normal.rs:83-121,187-253, computes the same unique value
// a*b*2^-256 mod p in [0,p)).
//
// The chain primitives have two bodies: PTX (device) and portable C (host).  The host bodies
// exist so tests/host_emu.cpp can run the very same pipeline logic on the CPU; they are also
// what the host side of the product uses to add the <= 8 per-GPU partial sums.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KGR_HD __host__ __device__ __forceinline__
#define KGR_D __device__ __forceinline__
#else
#define KGR_HD inline
#define KGR_D inline
#endif

namespace kgr {

// ---- field parameter packs (limbs little endian, 32-bit) -----------------------------------
// bn254/src/fq.rs:10-44
struct FqP {
    static KGR_HD constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static KGR_HD constexpr uint32_t one(int i) {  // R = 2^256 mod q
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static KGR_HD constexpr uint32_t r2(int i) {  // R^2 mod q
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static KGR_HD constexpr uint32_t r3(int i) {  // R^3 mod q (fq.rs:36-41)
        constexpr uint32_t m[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;  // low word of fq.rs:44
};
// bn254/src/fr.rs:11-51
struct FrP {
    static KGR_HD constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static KGR_HD constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static KGR_HD constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static KGR_HD constexpr uint32_t r3(int i) {  // R^3 mod r (fr.rs:43-48)
        constexpr uint32_t m[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;  // low word of fr.rs:51
};

template <class P> struct Fp {
    uint32_t v[8];
};

// ---- carry-chain primitives ----------------------------------------------------------------
// acc(8 limbs) += {a[0],a[2],a[4],a[6]} * b laid out at limbs (0,1)(2,3)(4,5)(6,7); the carry out
// of limb 7 is added to `top`.
KGR_HD void chain_cmad(uint32_t acc[8], uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b, uint32_t &top) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
    const uint32_t a[4] = {a0, a2, a4, a6};
    uint64_t c = 0;
    for (int k = 0; k < 4; k++) {
        uint64_t pr = (uint64_t)a[k] * b;
        uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)pr + c;
        acc[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(pr >> 32) + (lo >> 32);
        acc[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
    top += (uint32_t)c;
#endif
}

// Same as chain_cmad but the carry out of limb 7 is known to be zero (see fp_mul) and dropped.
KGR_HD void chain_cmad_nc(uint32_t acc[8], uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
    uint32_t top = 0;
    chain_cmad(acc, a0, a2, a4, a6, b, top);
#endif
}

// Shift-and-accumulate for the accumulator that changes alignment this round:
//   e0 += o[1] (carry c);  o[j] = {a1,a3,a5,a7}*b + o[j+2] + c  for the four 64-bit slots, o[8..9] = 0.
KGR_HD void chain_madc_rshift(uint32_t o[8], uint32_t &e0, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
        "madc.hi.u32 %7, %12, %13, 0;"
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(e0)
        : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
#else
    const uint32_t a[4] = {a1, a3, a5, a7};
    uint64_t s = (uint64_t)e0 + o[1];
    e0 = (uint32_t)s;
    uint64_t c = s >> 32;
    for (int k = 0; k < 4; k++) {
        uint64_t pr = (uint64_t)a[k] * b;
        uint32_t in_lo = (2 * k + 2 < 8) ? o[2 * k + 2] : 0u;
        uint32_t in_hi = (2 * k + 3 < 8) ? o[2 * k + 3] : 0u;
        uint64_t lo = (uint64_t)in_lo + (uint32_t)pr + c;
        uint64_t hi = (uint64_t)in_hi + (uint32_t)(pr >> 32) + (lo >> 32);
        o[2 * k] = (uint32_t)lo;
        o[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
#endif
}

// r = a + b (8 limbs), returns nothing: callers guarantee no carry out of limb 7.
KGR_HD void chain_add8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t s = (uint64_t)a[i] + b[i] + c;
        r[i] = (uint32_t)s;
        c = s >> 32;
    }
#endif
}

// r = a - b (8 limbs); returns the borrow as an all-ones / zero mask.
KGR_HD uint32_t chain_sub8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t mask;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(mask)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t brw = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - brw;
        r[i] = (uint32_t)d;
        brw = (d >> 32) & 1;
    }
    mask = (uint32_t)(0u - (uint32_t)brw);
#endif
    return mask;
}

template <class P> KGR_HD void load_mod(uint32_t m[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::mod(i);
}

// ---- field operations ----------------------------------------------------------------------
template <class P> KGR_HD Fp<P> fp_zero() {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
template <class P> KGR_HD Fp<P> fp_one() {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
}
template <class P> KGR_HD bool fp_is_zero(const Fp<P> &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
template <class P> KGR_HD bool fp_eq(const Fp<P> &a, const Fp<P> &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}

// if (x >= p) x -= p, for x < 2p  (normal.rs:13-30 tail)
template <class P> KGR_HD void fp_final_sub(uint32_t x[8]) {
    uint32_t m[8], t[8];
    load_mod<P>(m);
    uint32_t borrow = chain_sub8(t, x, m);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

// normal.rs:4-31
template <class P> KGR_HD Fp<P> fp_add(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    chain_add8(r.v, a.v, b.v);  // a + b < 2p < 2^255: no carry out
    fp_final_sub<P>(r.v);
    return r;
}
// normal.rs:34-53
template <class P> KGR_HD Fp<P> fp_sub(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    uint32_t d[8], m[8];
    uint32_t borrow = chain_sub8(d, a.v, b.v);
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::mod(i) & borrow;
    chain_add8(r.v, d, m);
    return r;
}
// normal.rs:56-80
template <class P> KGR_HD Fp<P> fp_dbl(const Fp<P> &a) { return fp_add(a, a); }
// normal.rs:170-184
template <class P> KGR_HD Fp<P> fp_neg(const Fp<P> &a) {
    Fp<P> r;
    uint32_t m[8];
    load_mod<P>(m);
    (void)chain_sub8(r.v, m, a.v);
    bool z = fp_is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
    return r;
}
// sign ? -a : a
template <class P> KGR_HD Fp<P> fp_cneg(const Fp<P> &a, bool sign) {
    Fp<P> n = fp_neg(a), r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = sign ? n.v[i] : a.v[i];
    return r;
}

// Montgomery product a*b*2^-256 mod p, fully reduced.
//
// Running total T = EV + 2^32*OD with EV, OD eight limbs each.  Round i adds a*b_i and m*p
// (m = T mod 2^32 * INV) so that T becomes divisible by 2^32; dividing by 2^32 swaps the roles of
// the two accumulators (the odd-aligned one becomes even-aligned), and the accumulator that was
// even-aligned is shifted down one 64-bit slot while the next row is added (chain_madc_rshift),
// its orphan limb 1 going into limb 0 of the other.  Bounds: a, b < p < 2^254 keep T < 2^287
// before each division, hence OD < 2^256 (no carry out of the odd chain) and the carry out of the
// even chain fits in od[7].  Final value < 2p, one conditional subtraction.
template <class P> KGR_HD Fp<P> fp_mul(const Fp<P> &a, const Fp<P> &b) {
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__) && !defined(KGR_HOST_EMULATE_CHAINS)
    // Host side of the product (combining per-GPU partial sums, window Horner): word-serial CIOS
    // on 4 x u64 with unsigned __int128.  tests/host_emu.cpp defines KGR_HOST_EMULATE_CHAINS to
    // exercise the even/odd chain algorithm below instead.
    typedef unsigned __int128 u128;
    uint64_t x[4], y[4], m[4], t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        x[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
        y[i] = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
        m[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
    }
    // -p^-1 mod 2^64 from the 32-bit constant by one Newton step: inv64 = inv32 * (2 + p0 * inv32)
    uint64_t inv = (uint64_t)P::INV;
    inv *= 2 + m[0] * inv;
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)x[j] * y[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * inv;
        c = ((u128)k * m[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)k * m[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    uint64_t d[4], brw = 0;
    for (int i = 0; i < 4; i++) {
        u128 s = (u128)t[i] - m[i] - brw;
        d[i] = (uint64_t)s;
        brw = (uint64_t)(s >> 64) & 1;
    }
    bool ge = t[4] != 0 || brw == 0;
    Fp<P> r;
    for (int i = 0; i < 4; i++) {
        uint64_t v = ge ? d[i] : t[i];
        r.v[2 * i] = (uint32_t)v;
        r.v[2 * i + 1] = (uint32_t)(v >> 32);
    }
    return r;
#else
    uint32_t ev[8], od[8];
    const uint32_t *x = a.v;
    // round 0: plain products
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint64_t pe = (uint64_t)x[2 * k] * b.v[0];
        uint64_t po = (uint64_t)x[2 * k + 1] * b.v[0];
        ev[2 * k] = (uint32_t)pe;
        ev[2 * k + 1] = (uint32_t)(pe >> 32);
        od[2 * k] = (uint32_t)po;
        od[2 * k + 1] = (uint32_t)(po >> 32);
    }
    uint32_t m = ev[0] * P::INV;
    chain_cmad_nc(od, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
    chain_cmad(ev, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, od[7]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        // after the division by 2^32: E = previous odd accumulator, O = previous even accumulator
        uint32_t *E = (i & 1) ? od : ev;
        uint32_t *O = (i & 1) ? ev : od;
        chain_madc_rshift(O, E[0], x[1], x[3], x[5], x[7], b.v[i]);
        chain_cmad(E, x[0], x[2], x[4], x[6], b.v[i], O[7]);
        m = E[0] * P::INV;
        chain_cmad_nc(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        chain_cmad(E, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, O[7]);
    }
    // round 7 ran with E = od (so od[0] == 0 now) and O = ev: result = T / 2^32 = ev + od / 2^32
    Fp<P> r;
    uint32_t sh[8];
#pragma unroll
    for (int i = 0; i < 7; i++) sh[i] = od[i + 1];
    sh[7] = 0;
    chain_add8(r.v, ev, sh);
    fp_final_sub<P>(r.v);
    return r;
#endif
}

template <class P> KGR_HD Fp<P> fp_sqr(const Fp<P> &a) { return fp_mul(a, a); }

// fq.rs:94-100 / fr.rs:122-128: Montgomery -> canonical (multiply by the integer 1)
template <class P> KGR_HD Fp<P> fp_from_mont(const Fp<P> &a) {
    Fp<P> one_int = fp_zero<P>();
    one_int.v[0] = 1;
    return fp_mul(a, one_int);
}
template <class P> KGR_HD Fp<P> fp_to_mont(const Fp<P> &a) {
    Fp<P> r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
    return fp_mul(a, r2);
}

// represent.rs:18-28: (lo*R^2 + hi*R^3) Montgomery products = (lo + 2^256 hi) mod p in Montgomery form
template <class P> KGR_HD Fp<P> fp_from_u512(const uint32_t w[16]) {
    Fp<P> lo, hi, r2, r3;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lo.v[i] = w[i];
        hi.v[i] = w[8 + i];
        r2.v[i] = P::r2(i);
        r3.v[i] = P::r3(i);
    }
    return fp_add(fp_mul(lo, r2), fp_mul(hi, r3));
}

// a^(p-2) (normal.rs:256-287); zero maps to zero (callers test for zero where the reference returns None)
template <class P> KGR_HD Fp<P> fp_inv(const Fp<P> &a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::mod(i);
    e[0] -= 2;  // both moduli have low limb >= 2
    Fp<P> acc = fp_one<P>();
    for (int i = 255; i >= 0; i--) {
        acc = fp_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fp_mul(acc, a);
    }
    return acc;
}

// ---- Fq2 = Fq[u] / (u^2 + 1)  (bn254/src/fqn.rs:347-370; limbs.rs tower macros) -------------
// The coordinates of G2.  Products use three base-field multiplications (Karatsuba) and squares two
// ((a0+a1)(a0-a1), 2 a0 a1) where the reference spends four and three: same field element, fewer multiplier cycles.
template <class P> struct Fp2 {
    Fp<P> c0, c1;
};
template <class P> KGR_HD bool fp_is_zero(const Fp2<P> &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
template <class P> KGR_HD bool fp_eq(const Fp2<P> &a, const Fp2<P> &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
template <class P> KGR_HD Fp2<P> fp_add(const Fp2<P> &a, const Fp2<P> &b) { return Fp2<P>{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
template <class P> KGR_HD Fp2<P> fp_sub(const Fp2<P> &a, const Fp2<P> &b) { return Fp2<P>{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
template <class P> KGR_HD Fp2<P> fp_dbl(const Fp2<P> &a) { return Fp2<P>{fp_dbl(a.c0), fp_dbl(a.c1)}; }
template <class P> KGR_HD Fp2<P> fp_neg(const Fp2<P> &a) { return Fp2<P>{fp_neg(a.c0), fp_neg(a.c1)}; }
template <class P> KGR_HD Fp2<P> fp_cneg(const Fp2<P> &a, bool sign) { return Fp2<P>{fp_cneg(a.c0, sign), fp_cneg(a.c1, sign)}; }
// fqn.rs:359-363
template <class P> KGR_HD Fp2<P> fp2_mul_inline(const Fp2<P> &a, const Fp2<P> &b) {
    Fp<P> t0 = fp_mul(a.c0, b.c0);
    Fp<P> t1 = fp_mul(a.c1, b.c1);
    Fp<P> t2 = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
    return Fp2<P>{fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1)};
}
// fqn.rs:365-369
template <class P> KGR_HD Fp2<P> fp2_sqr_inline(const Fp2<P> &a) {
    Fp<P> t = fp_mul(a.c0, a.c1);
    return Fp2<P>{fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1)), fp_dbl(t)};
}
#if defined(__CUDA_ARCH__) && defined(KGR_FP2_CALLS)
// One shared copy of the Fq2 product / square per kernel instead of ~30 inlined ones per point addition: the inlined G2 accumulate
// loop is > 100 KB of SASS and stalls on instruction fetch (profiles/r01_g2.md).
template <class P> __device__ __noinline__ Fp2<P> fp2_mul_call(Fp2<P> a, Fp2<P> b) { return fp2_mul_inline(a, b); }
template <class P> __device__ __noinline__ Fp2<P> fp2_sqr_call(Fp2<P> a) { return fp2_sqr_inline(a); }
template <class P> KGR_HD Fp2<P> fp_mul(const Fp2<P> &a, const Fp2<P> &b) { return fp2_mul_call(a, b); }
template <class P> KGR_HD Fp2<P> fp_sqr(const Fp2<P> &a) { return fp2_sqr_call(a); }
#else
template <class P> KGR_HD Fp2<P> fp_mul(const Fp2<P> &a, const Fp2<P> &b) { return fp2_mul_inline(a, b); }
template <class P> KGR_HD Fp2<P> fp_sqr(const Fp2<P> &a) { return fp2_sqr_inline(a); }
#endif
// fqn.rs:348-357: conj(a) / (a0^2 + a1^2); zero maps to zero
template <class P> KGR_HD Fp2<P> fp_inv(const Fp2<P> &a) {
    Fp<P> t = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
    return Fp2<P>{fp_mul(t, a.c0), fp_mul(t, fp_neg(a.c1))};
}

// Element traits so the group law and the MSM pipeline are written once for Fp and Fp2 coordinates.
template <class E> struct El;
template <class P> struct El<Fp<P>> {
    static constexpr int WORDS = 8;
    static KGR_HD Fp<P> zero() { return fp_zero<P>(); }
    static KGR_HD Fp<P> one() { return fp_one<P>(); }
    static KGR_HD uint32_t &word(Fp<P> &a, int i) { return a.v[i]; }
    static KGR_HD uint32_t word(const Fp<P> &a, int i) { return a.v[i]; }
};
template <class P> struct El<Fp2<P>> {
    static constexpr int WORDS = 16;
    static KGR_HD Fp2<P> zero() { return Fp2<P>{fp_zero<P>(), fp_zero<P>()}; }
    static KGR_HD Fp2<P> one() { return Fp2<P>{fp_one<P>(), fp_zero<P>()}; }
    static KGR_HD uint32_t &word(Fp2<P> &a, int i) { return i < 8 ? a.c0.v[i] : a.c1.v[i - 8]; }
    static KGR_HD uint32_t word(const Fp2<P> &a, int i) { return i < 8 ? a.c0.v[i] : a.c1.v[i - 8]; }
};

// 16-byte loads / stores of whole field elements (8 words, or 16 for Fq2) between global memory and registers
template <class E> KGR_HD void el_load(E &e, const void *src) {
#if defined(__CUDA_ARCH__)
    const uint4 *q = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int k = 0; k < El<E>::WORDS / 4; k++) {
        uint4 a = __ldg(q + k);
        El<E>::word(e, 4 * k) = a.x; El<E>::word(e, 4 * k + 1) = a.y; El<E>::word(e, 4 * k + 2) = a.z; El<E>::word(e, 4 * k + 3) = a.w;
    }
#else
    e = *reinterpret_cast<const E *>(src);
#endif
}
template <class E> KGR_HD void el_store(void *dst, const E &e) {
#if defined(__CUDA_ARCH__)
    uint4 *q = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int k = 0; k < El<E>::WORDS / 4; k++)
        q[k] = make_uint4(El<E>::word(e, 4 * k), El<E>::word(e, 4 * k + 1), El<E>::word(e, 4 * k + 2), El<E>::word(e, 4 * k + 3));
#else
    *reinterpret_cast<E *>(dst) = e;
#endif
}

}  // namespace kgr
