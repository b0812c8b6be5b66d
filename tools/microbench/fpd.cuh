// Fp on the FP64 pipe (BASELINE.json:5 hot path, K1): B200 issues 64 DFMA/clk/SM on a pipe that is idle while
// the integer multiplier of field.cuh saturates the IMAD pipe (32 IMAD.WIDE/clk/SM), so a second
// representation of the same field lets both pipes work on different points at once.
//
// Representation: x = sum v[i] * 2^(48 i), v[i] an integer in [0, 2^48) held exactly in a double, x in [0, 2p).
// Same Montgomery radix as field.cuh (R = 2^384 = 2^(48*8) = 2^(32*12)): converting is a re-slicing of bits.
//
// Product of two limbs (< 2^96) through FMA, all operations exact:
//   h' = fma_rz(a, b, h)        h in [2^100, 2^101) has ulp 2^48, so h' - h = floor(a*b / 2^48) * 2^48
//   lo = fma(a, b, -(h' - h))   the low 48 bits, exact
// The h chain of one column accumulates up to 16 high halves (16 * (2^48 - 2) * 2^48 < 2^100) before it is
// unbiased; low halves are summed as doubles (< 30 * 2^48 < 2^53 per column, see fpd_mul).
// R > 4p, so operands below 2p give a result below 2p with no conditional subtraction ("lazy" Montgomery).
#pragma once
#include "../../kzg_batch_verification_scheme_b200/csrc/field.cuh"

struct FpD { double v[8]; };

#if defined(KZGB_EMU)
// host emulation (tests only): the caller runs under fesetround(FE_TOWARDZERO); every operation other than the
// h chain is exact, hence independent of the rounding mode.
#include <cmath>
#define FPD_FMA_RZ(a, b, c) std::fma((a), (b), (c))
#define FPD_FMA(a, b, c) std::fma((a), (b), (c))
#define FPD_ADD(a, b) ((a) + (b))
#define FPD_MUL(a, b) ((a) * (b))
#define FPD_U2D(x) ((double)(x))
#define FPD_D2U(x) ((u64)(x))
#else
#define FPD_FMA_RZ(a, b, c) __fma_rz((a), (b), (c))
#define FPD_FMA(a, b, c) __fma_rn((a), (b), (c))
#define FPD_ADD(a, b) __dadd_rn((a), (b))
#define FPD_MUL(a, b) __dmul_rn((a), (b))
#define FPD_U2D(x) __ull2double_rn(x)
#define FPD_D2U(x) __double2ull_rz(x)
#endif

#define FPD_C100 1267650600228229401496703205376.0   /* 2^100 */
#define FPD_2P52 4503599627370496.0                  /* 2^52  */
#define FPD_2P48 281474976710656.0                   /* 2^48  */
#define FPD_2M48 3.5527136788005009e-15              /* 2^-48 */

// one limb product into column accumulators: Lk += low 48 bits, Hk chain += high part (biased by 2^100)
#define FPD_PROD(x, y, Lk, Hk)                          \
    {                                                   \
        double hn_ = FPD_FMA_RZ((x), (y), (Hk));        \
        double d_ = FPD_ADD(hn_, -(Hk));                \
        (Lk) = FPD_ADD((Lk), FPD_FMA((x), (y), -d_));   \
        (Hk) = hn_;                                     \
    }
// high-part chain -> integer count of 2^48 units
#define FPD_UNBIAS(h) FPD_FMA((h), FPD_2M48, -FPD_2P52)
// floor(t / 2^48) for 0 <= t < 2^53
#define FPD_FLOOR48(t) FPD_ADD(FPD_FMA_RZ((t), FPD_2M48, FPD_2P52), -FPD_2P52)

// Montgomery reduction of the 16 column sums (L, H from the a*b part) interleaved with the q*p products.
KZ_HD FpD fpd_redc(double* L, double* H) {
    FpD r;
    double carry = 0.0;
    KZ_UNROLL
    for (int k = 0; k < 8; ++k) {
        double t = FPD_ADD(L[k], carry);
        if (k) t = FPD_ADD(t, FPD_UNBIAS(H[k - 1]));
        double cf = FPD_FLOOR48(t);
        double tl = FPD_FMA(cf, -FPD_2P48, t);                       // t mod 2^48
        double hq = FPD_FMA_RZ(tl, FPD_M48, FPD_C100);
        double q = FPD_FMA(tl, FPD_M48, -FPD_ADD(hq, -FPD_C100));    // q = tl * (-1/p) mod 2^48
        {
            double hn = FPD_FMA_RZ(q, FPD_P[0], H[k]);
            double d = FPD_ADD(hn, -H[k]);
            double lo0 = FPD_FMA(q, FPD_P[0], -d);
            H[k] = hn;
            carry = FPD_MUL(FPD_ADD(t, lo0), FPD_2M48);              // column k is now 0 mod 2^48
        }
        KZ_UNROLL
        for (int j = 1; j < 8; ++j) FPD_PROD(q, FPD_P[j], L[k + j], H[k + j]);
    }
    KZ_UNROLL
    for (int k = 8; k < 16; ++k) {
        double t = FPD_ADD(FPD_ADD(L[k], carry), FPD_UNBIAS(H[k - 1]));
        double cf = FPD_FLOOR48(t);
        r.v[k - 8] = FPD_FMA(cf, -FPD_2P48, t);
        carry = cf;
    }
    return r;
}

KZ_HD FpD fpd_mul(const FpD& a, const FpD& b) {
    double L[16], H[16];
    KZ_UNROLL
    for (int k = 0; k < 16; ++k) { L[k] = 0.0; H[k] = FPD_C100; }
    KZ_UNROLL
    for (int i = 0; i < 8; ++i) {
        KZ_UNROLL
        for (int j = 0; j < 8; ++j) FPD_PROD(a.v[i], b.v[j], L[i + j], H[i + j]);
    }
    return fpd_redc(L, H);
}

KZ_HD FpD fpd_sqr(const FpD& a) {
    double L[16], H[16], a2[8];
    KZ_UNROLL
    for (int k = 0; k < 16; ++k) { L[k] = 0.0; H[k] = FPD_C100; }
    KZ_UNROLL
    for (int i = 0; i < 8; ++i) a2[i] = FPD_ADD(a.v[i], a.v[i]);
    // cross products with one operand doubled (< 2^97: a high half counts twice in the chain budget, 4 per
    // column at most, so the a^2 part uses at most 8 of the 16 slots like a general product)
    KZ_UNROLL
    for (int i = 0; i < 8; ++i) {
        FPD_PROD(a.v[i], a.v[i], L[2 * i], H[2 * i]);
        KZ_UNROLL
        for (int j = i + 1; j < 8; ++j) FPD_PROD(a2[i], a.v[j], L[i + j], H[i + j]);
    }
    return fpd_redc(L, H);
}

// ---- conversions (values stay in Montgomery form; only the limb width changes)
KZ_HD FpD fpd_from_fp(const Fp& a) {
    FpD r;
    KZ_UNROLL
    for (int k = 0; k < 4; ++k) {
        u64 w0 = a.v[3 * k], w1 = a.v[3 * k + 1], w2 = a.v[3 * k + 2];
        r.v[2 * k] = FPD_U2D(w0 | ((w1 & 0xFFFFull) << 32));
        r.v[2 * k + 1] = FPD_U2D((w1 >> 16) | (w2 << 16));
    }
    return r;
}
// to canonical 12 x 32 limbs; the input may be anywhere in [0, 2p)
KZ_HD Fp fpd_to_fp(const FpD& a) {
    Fp r;
    KZ_UNROLL
    for (int k = 0; k < 4; ++k) {
        u64 e = FPD_D2U(a.v[2 * k]), o = FPD_D2U(a.v[2 * k + 1]);
        r.v[3 * k] = (u32)e;
        r.v[3 * k + 1] = (u32)(e >> 32) | ((u32)o << 16);
        r.v[3 * k + 2] = (u32)(o >> 16);
    }
#if defined(KZGB_EMU)
    return fp_add(r, fp_zero());          // host: one conditional subtraction through the portable adder
#else
    return fp_reduce_once(r);
#endif
}
