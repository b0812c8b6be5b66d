// Pairing check without the Horner tail (device bodies; kernels in k_mpair.cu).
//
// The MSM stage ends with 129 slice sums U_0 .. U_128 per sum (msm.cuh: U_j = sum of the buckets whose magnitude has
// bit j - c*w set, or the bucket of magnitude 2^k of window w), and the value of the sum is  S = sum_j 2^j U_j.
// Finishing that on the G1 side is a serial chain of ~128 doublings (0.9 ms on one thread, 0.4 ms on a quad).  By
// bilinearity the doublings can be pushed onto the FIXED G2 arguments instead:
//
//     e(sum_t 2^(4t) V_t, Q) = prod_t e(V_t, [2^(4t)] Q),      V_t = sum_{u<4} 2^u U_(4t+u)     (33 terms per sum)
//
// so the check  e(A, G2) e(B, [tau]G2) = 1  becomes a product of 66 pairings against 66 fixed G2 points whose line
// coefficients are precomputed once per context (33 multiples of G2 and of [tau]G2; 0.86 MB).  Nothing on the G1
// side is serial any more:
//   k_mp_terms   3 short Horner chains per term (3 doublings + 4 additions, one quad each)          ~25 product rounds
//   k_mp_coefs   sum of the shards' terms, then (ZZ ZZZ, X ZZZ, Y ZZ) per pair                        1 round (+ 4 per extra shard)
//   k_mp_lines   for every Miller step the 66 line values and their product (a tree of Fp12 products, one block per
//                (step, group), all SMs busy)                                                      ~10 product rounds of latency
//   k_mp_merge   per Miller iteration: product of the groups' partial products (doubling and addition step)
//   k_mp_check   the only serial kernel: f <- f^2 F_i over 63 iterations, then the final exponentiation as an
//                inversion-free equality test (below)
//
// Final exponentiation without an inversion.  With u = |x|, m = f^(p^2+1) and 3(p^4-p^2+1)/r = H+ - H-,
//     H+ = (u+1)^2 (p u^2 + p^3 + u) + 3,      H- = (u+1)^2 (p + u^3 + u p^2)
// (tools/check_final_exp.py), f^(3(p^12-1)/r) = conj(h)/h for h = m^(H+) / m^(H-), so the product of pairings is one iff
//     conj(X+) X-  ==  X+ conj(X-),        X+- = m^(H+-),
// five exponentiations by u on general (non-unitary) elements and no Fp inversion at all.
//
// One Fp12 product = one "round": 108 threads each do ONE Montgomery product (Karatsuba parts of the 36 Fp2 partial
// products), one barrier, then 96 threads fold them (12 outputs x 8 lanes, xor-shuffle tree).  Measured latencies
// (profiles/r2_fp_latency.txt): product 1792 clk, modular add 80 clk, barrier 21 clk.
#pragma once
#include "msm.cuh"
#include "pairing.cuh"
#include "quad.cuh"

#define KZ_MP_G 4                                    // bit positions per term
#define KZ_MP_TERMS 33                               // ceil(129 / KZ_MP_G)
#define KZ_MP_PAIRS (2 * KZ_MP_TERMS)                // A-side terms, then B-side terms
#define KZ_MP_ITERS 63                               // Miller iterations (bits 62..0 of |x|)

struct MpCoef { Fp alpha, beta, gamma; u32 inf, pad[3]; };      // line at P: a*alpha + (b*beta) w^2 + gamma w^3
struct MpSumDesc { const G1Xyzz* slices; const G1Xyzz* buckets; int c, W, nbits; };     // slices == null: the zero sum

// U_j of a sum, j = absolute bit position 0 .. nbits
KZ_HD G1Xyzz mp_slice(const MpSumDesc& d, int j) {
    if (!d.slices || j > d.nbits) return xyzz_inf();
    const int signed_bits = d.c * (d.W - 1);
    if (j == d.nbits) {                              // magnitude 2^tb of the unsigned top window = the last bucket
        const int tb = d.nbits - signed_bits;
        return d.buckets[((size_t)(d.W - 1) << (d.c - 1)) + ((size_t)1 << tb) - 1];
    }
    if (j < signed_bits && (j % d.c) == d.c - 1)     // magnitude 2^(c-1) of a signed window
        return d.buckets[(((size_t)(j / d.c) + 1) << (d.c - 1)) - 1];
    return d.slices[j];
}
// V_t = sum_{u<4} 2^u U_(4t+u) by Horner (quad arithmetic)
KZ_HD G1Xyzz mp_term(Quad& q, const MpSumDesc& d, int t) {
    G1Xyzz acc = xyzz_inf();
    for (int u = KZ_MP_G - 1; u >= 0; --u) {
        acc = quad_xyzz_dbl(q, acc);
        acc = quad_xyzz_add(q, acc, mp_slice(d, KZ_MP_G * t + u));
    }
    return acc;
}
// (ZZ ZZZ, X ZZZ, Y ZZ) of a term: the three products every line of its pair needs.  With x = X/ZZ, y = Y/ZZZ the line
// a + (b x) w^2 + y w^3 scaled by ZZ ZZZ (a factor in Fp, killed by the final exponentiation) is a alpha + (b beta) w^2 + gamma w^3.
KZ_HD MpCoef mp_coef_of(Quad& q, const G1Xyzz& acc) {
    MpCoef c;
    Fp d0;
    quad_mul4(q, acc.ZZ, acc.ZZZ, acc.X, acc.ZZZ, acc.Y, acc.ZZ, acc.Y, acc.ZZ, c.alpha, c.beta, c.gamma, d0);
    c.inf = xyzz_is_inf(acc) ? 1u : 0u;
    c.pad[0] = c.pad[1] = c.pad[2] = 0;
    return c;
}
// Fp coefficient ci (0..11: c[ci/2].c0 / .c1) of the line value of Miller step s at a pair, as a dense Fp12;
// a pair at infinity contributes the constant 1
KZ_HD Fp mp_line_coeff(const G2Lines& T, int s, const MpCoef& cf, int ci) {
    const bool inf = cf.inf != 0;
    if (!inf && (ci < 2 || ci == 4 || ci == 5)) {
        const Fp2& src = ci < 2 ? T.a[s] : T.b[s];
        return fp_mul((ci & 1) ? src.c1 : src.c0, ci < 2 ? cf.alpha : cf.beta);
    }
    if (!inf && ci == 6) return cf.gamma;
    if (inf && ci == 0) return fp_one();
    return fp_zero();
}
// Miller step index of iteration it (bit 62 - it): doubling step, and whether an addition step follows it
KZ_HD int mp_step_of_iter(int it, bool& has_add) {
    const u64 k = ((u64)X_ABS_HI << 32) | X_ABS_LO;
    int s = 0;
    for (int i = 62; i > 62 - it; --i) s += 1 + (int)((k >> i) & 1);
    has_add = (k >> (62 - it)) & 1;
    return s;
}

// ------------------------------------------------------------------ one Fp12 product by a unit of 128 threads
// Measured per round on an idle SM (tools/microbench/mp_round.cu): products 1.9-2.3k clk, fold 0.8-1.1k clk, barriers 10 clk.
struct MpUnit { Fp kar[108]; Fp zero; };             // Karatsuba parts of one Fp12 product; a zero operand
// phase 1 (t < 108): Karatsuba parts v0 = a0 b0, v1 = a1 b1, v2 = (a0+a1)(b0+b1) of the 36 Fp2 partial products.
// Uniform code: every lane adds two operands, the second one being the zero constant for v0 and v1.
KZ_HD void mp_mul_products(MpUnit& U, int t, const Fp12& x, const Fp12& y) {
    if (t < 108) {
        const int q = t / 3, part = t - 3 * q, i = q / 6, j = q - 6 * i;
        const Fp* a1 = part == 1 ? &x.c[i].c1 : &x.c[i].c0;
        const Fp* a2 = part == 2 ? &x.c[i].c1 : &U.zero;
        const Fp* b1 = part == 1 ? &y.c[j].c1 : &y.c[j].c0;
        const Fp* b2 = part == 2 ? &y.c[j].c1 : &U.zero;
        U.kar[t] = fp_mul(fp_add(*a1, *a2), fp_add(*b1, *b2));
    }
}
// phase 2 (after a barrier): coefficient k, real (h = 0) / imaginary (h = 1) part; lane i < 6 holds the contribution of
// the partial product (i, j) with i + j = k (mod 6):
//   i + j = k      real  v0 - v1            imaginary  v2 - v1 - v0
//   i + j = k + 6  real  (v0 - v2) + v0     imaginary  v2 - v1 - v1       (times xi = 1 + u)
KZ_HD Fp mp_fold_contrib(const MpUnit& U, int k, int h, int i) {
    int j = k - i;
    const bool hi = j < 0;
    if (hi) j += 6;
    const int q = 3 * (i * 6 + j);
    if (h == 0) {
        const Fp X = U.kar[q], Y = U.kar[q + (hi ? 2 : 1)];
        const Fp d = fp_sub(X, Y), dp = fp_add(d, X);
        return hi ? dp : d;
    }
    const Fp X = U.kar[q + 2], Y = U.kar[q + 1], T = U.kar[q + (hi ? 1 : 0)];
    return fp_sub(fp_sub(X, Y), T);
}
#if defined(KZGB_EMU)
#define MP_TID() 0
KZ_HD void mp_unit_init(MpUnit& U) { U.zero = fp_zero(); }
KZ_HD void mp_mul(MpUnit& U, Fp12& dst, const Fp12& x, const Fp12& y) {
    for (int t = 0; t < 108; ++t) mp_mul_products(U, t, x, y);
    for (int o = 0; o < 12; ++o) {
        Fp r = fp_zero();
        for (int i = 0; i < 6; ++i) r = fp_add(r, mp_fold_contrib(U, o >> 1, o & 1, i));
        if (o & 1) dst.c[o >> 1].c1 = r; else dst.c[o >> 1].c0 = r;
    }
}
KZ_HD void mp_mul_cold(MpUnit& U, Fp12& dst, const Fp12& x, const Fp12& y) { mp_mul(U, dst, x, y); }
#else
#define MP_TID() ((int)(threadIdx.x & 127u))
KZ_HD void mp_unit_init(MpUnit& U) { if (MP_TID() == 0) U.zero = fp_zero(); }       // followed by a barrier of the caller
// Four warps: warps 0, 1 fold the real parts (warp-uniform arithmetic), warps 2, 3 the imaginary parts; 3 outputs per
// warp x 8 lanes, xor-shuffle tree over the 6 contributions.  dst may alias the operands of phase 1.
KZ_HD void mp_mul_fold(const MpUnit& U, int t, Fp12& dst) {
    const int w = t >> 5, l = t & 31, g = l >> 3, i = l & 7;
    const int h = w >> 1, k = (w & 1) * 3 + (g < 3 ? g : 0);
    Fp r = (g < 3 && i < 6) ? mp_fold_contrib(U, k, h, i) : fp_zero();
    KZ_UNROLL for (int s = 4; s; s >>= 1) {
        Fp ot;
        KZ_UNROLL for (int m = 0; m < 12; ++m) ot.v[m] = __shfl_xor_sync(0xFFFFFFFFu, r.v[m], s);
        r = fp_add(r, ot);
    }
    if (g < 3 && i == 0) { if (h) dst.c[k].c1 = r; else dst.c[k].c0 = r; }
}
// whole product by a block that IS one unit (128 threads)
KZ_HD void mp_mul(MpUnit& U, Fp12& dst, const Fp12& x, const Fp12& y) {
    const int t = MP_TID();
    mp_mul_products(U, t, x, y);
    __syncthreads();
    mp_mul_fold(U, t, dst);
    __syncthreads();
}
KZ_COLD void mp_mul_cold(MpUnit& U, Fp12& dst, const Fp12& x, const Fp12& y) { mp_mul(U, dst, x, y); }
#endif

// ------------------------------------------------------------------ the serial part (one unit)
struct MpScratch {
    MpUnit U;
    Fp12 f, fb, m, m1, m2, n, n1, n2, n3, t, a, b;
    int result;
};
// dst = x^u, u = |x| = 0xd201000000010000 (63 squarings, 5 products); dst must not alias x; uses S.t
KZ_COLD void mp_pow_u(MpScratch& S, Fp12& dst, const Fp12& x) {
    const u64 k = ((u64)X_ABS_HI << 32) | X_ABS_LO;
    coop_copy(S.t, x);
    for (int i = 62; i >= 0; --i) {
        mp_mul(S.U, S.t, S.t, S.t);
        if ((k >> i) & 1) mp_mul(S.U, S.t, S.t, x);
    }
    coop_copy(dst, S.t);
}
// In: S.f = product of the Miller functions (already conjugated for x < 0).  Out: S.result = 1 iff f^((p^12-1)/r) == 1.
KZ_COLD void mp_final_check(MpScratch& S) {
    coop_frob2(S.a, S.f);
    mp_mul_cold(S.U, S.m, S.a, S.f);                  // m = f^(p^2+1)
    mp_pow_u(S, S.m1, S.m);
    mp_pow_u(S, S.m2, S.m1);
    mp_mul_cold(S.U, S.a, S.m1, S.m1);
    mp_mul_cold(S.U, S.a, S.a, S.m2);
    mp_mul_cold(S.U, S.n, S.a, S.m);                  // n = m^((u+1)^2)
    mp_pow_u(S, S.n1, S.n);
    mp_pow_u(S, S.n2, S.n1);
    mp_pow_u(S, S.n3, S.n2);
    // X+ = frob1(n2) * frob3(n) * n1 * m^3  -> S.m1
    mp_mul_cold(S.U, S.a, S.m, S.m);
    mp_mul_cold(S.U, S.a, S.a, S.m);                  // m^3
    mp_mul_cold(S.U, S.a, S.a, S.n1);
    coop_frob1(S.b, S.n2);
    mp_mul_cold(S.U, S.a, S.a, S.b);
    coop_frob2(S.b, S.n);
    coop_frob1(S.m2, S.b);                       // n^(p^3)
    mp_mul_cold(S.U, S.m1, S.a, S.m2);
    // X- = frob1(n) * n3 * frob2(n1)        -> S.m2
    coop_frob1(S.a, S.n);
    mp_mul_cold(S.U, S.a, S.a, S.n3);
    coop_frob2(S.b, S.n1);
    mp_mul_cold(S.U, S.m2, S.a, S.b);
    // conj(X+) X-  ==  X+ conj(X-)
    coop_conj(S.a, S.m1);
    mp_mul_cold(S.U, S.a, S.a, S.m2);
    coop_conj(S.b, S.m2);
    mp_mul_cold(S.U, S.b, S.b, S.m1);
    COOP_FOR(t, 1) {
        bool same = true;
        for (int k = 0; k < 6; ++k) same = same && fp2_eq(S.a.c[k], S.b.c[k]);
        S.result = same ? 1 : 0;
    }
    COOP_SYNC();
}
