// Quad-lane point arithmetic for the latency-bound tail of a batch (bucket reduction, subgroup chains, Horner).
//
// Measured on B200 (tools/microbench/fp_latency.cu, profiles/r2_fp_latency.txt): one warp issues a carry-chained wide
// multiply-add every ~6 clk no matter how many lanes are active or how much instruction-level parallelism the thread
// offers, so ONE Montgomery product costs 1792 clk of latency whether 1 or 32 lanes compute one; a modular addition is
// 80 clk, a shared-memory exchange between lanes ~80 clk.  A point operation executed by one thread is therefore 9-14
// products deep (xyzz_add 25.8k clk, xyzz_dbl 15.7k clk).  Here FOUR adjacent lanes (a "quad") execute one point
// operation: the independent products of each formula level go to different lanes of the same warp instruction
// (free), results are exchanged through 8 shared-memory slots per quad, and every lane redundantly evaluates the cheap
// linear steps.  XYZZ doubling becomes 3 product rounds (2/4/3 products), XYZZ addition 4 rounds (4/4/3/3).
//
// All lanes of a quad hold the same inputs and return the same outputs; control flow is uniform within a quad and may
// differ between the quads of a warp (the lane mask of every __syncwarp is the quad's).
// The host emulation (KZGB_EMU) evaluates the same formulas with the four products of a round computed in sequence.
#pragma once
#include "g1.cuh"

#if defined(KZGB_EMU)
struct Quad {};
KZ_HD Quad quad_make(void*, int) { return {}; }
KZ_HD void quad_mul4(Quad&, const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2, const Fp& a3,
                     const Fp& b3, Fp& r0, Fp& r1, Fp& r2, Fp& r3) {
    r0 = fp_mul(a0, b0); r1 = fp_mul(a1, b1); r2 = fp_mul(a2, b2); r3 = fp_mul(a3, b3);
}
#else
#define KZ_QUAD_SLOTS 8                    // two buffers of four product results
struct Quad {
    Fp* sm;                                // KZ_QUAD_SLOTS Fp of shared memory private to this quad
    u32 ql;                                // lane within the quad
    u32 mask;                              // the quad's lanes
    u32 phase;                             // buffer written by the next round
};
// sm_base: shared memory for all quads of the block, quad_in_block selects the region
KZ_HD Quad quad_make(Fp* sm_base, int quad_in_block) {
    Quad q;
    const u32 lane = threadIdx.x & 31u;
    q.sm = sm_base + (size_t)quad_in_block * KZ_QUAD_SLOTS;
    q.ql = lane & 3u;
    q.mask = 0xFu << (lane & 28u);
    q.phase = 0;
    return q;
}
KZ_HD Fp fp_sel4(u32 ql, const Fp& a0, const Fp& a1, const Fp& a2, const Fp& a3) {
    Fp r;
    const bool b0 = ql & 1u, b1 = ql & 2u;
    KZ_UNROLL for (int k = 0; k < 12; ++k) {
        u32 lo = b0 ? a1.v[k] : a0.v[k], hi = b0 ? a3.v[k] : a2.v[k];
        r.v[k] = b1 ? hi : lo;
    }
    return r;
}
// one product round: lane i of the quad computes a_i * b_i, afterwards every lane holds all four products.
// Double-buffered: the write of round k+2 into a buffer is separated from the reads of round k by the barrier of round k+1.
KZ_HD void quad_mul4(Quad& q, const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2, const Fp& a3,
                     const Fp& b3, Fp& r0, Fp& r1, Fp& r2, Fp& r3) {
    Fp a = fp_sel4(q.ql, a0, a1, a2, a3), b = fp_sel4(q.ql, b0, b1, b2, b3);
    Fp r = fp_mul(a, b);
    Fp* buf = q.sm + 4u * q.phase;
    q.phase ^= 1u;
    buf[q.ql] = r;
    __syncwarp(q.mask);
    r0 = buf[0]; r1 = buf[1]; r2 = buf[2]; r3 = buf[3];
}
#endif

// XYZZ doubling (dbl-2008-s-1), 3 rounds
KZ_HD G1Xyzz quad_xyzz_dbl(Quad& q, const G1Xyzz& p) {
    if (xyzz_is_inf(p)) return p;
    Fp U = fp_dbl(p.Y);
    Fp V, x2, d0, d1;
    quad_mul4(q, U, U, p.X, p.X, U, U, p.X, p.X, V, x2, d0, d1);
    Fp M = fp_add(fp_dbl(x2), x2);
    Fp W, S, MM, ZZ3;
    quad_mul4(q, U, V, p.X, V, M, M, V, p.ZZ, W, S, MM, ZZ3);
    G1Xyzz r;
    r.X = fp_sub(MM, fp_dbl(S));
    Fp t, u, ZZZ3;
    quad_mul4(q, M, fp_sub(S, r.X), W, p.Y, W, p.ZZZ, W, p.ZZZ, t, u, ZZZ3, d0);
    r.Y = fp_sub(t, u);
    r.ZZ = ZZ3;
    r.ZZZ = ZZZ3;
    return r;
}
// XYZZ + XYZZ (add-2008-s), 4 rounds; P + P, P - P and infinity handled exactly as xyzz_add does
KZ_HD G1Xyzz quad_xyzz_add(Quad& q, const G1Xyzz& p, const G1Xyzz& o) {
    if (xyzz_is_inf(p)) return o;
    if (xyzz_is_inf(o)) return p;
    Fp U1, U2, S1, S2;
    quad_mul4(q, p.X, o.ZZ, o.X, p.ZZ, p.Y, o.ZZZ, o.Y, p.ZZZ, U1, U2, S1, S2);
    Fp Pp = fp_sub(U2, U1), Rr = fp_sub(S2, S1);
    if (fp_is_zero(Pp)) return fp_is_zero(Rr) ? quad_xyzz_dbl(q, p) : xyzz_inf();
    Fp PP, RR, ZZ12, ZZZ12;
    quad_mul4(q, Pp, Pp, Rr, Rr, p.ZZ, o.ZZ, p.ZZZ, o.ZZZ, PP, RR, ZZ12, ZZZ12);
    Fp PPP, Q, ZZ3, d0;
    quad_mul4(q, Pp, PP, U1, PP, ZZ12, PP, ZZ12, PP, PPP, Q, ZZ3, d0);
    G1Xyzz r;
    r.X = fp_sub(fp_sub(RR, PPP), fp_dbl(Q));
    Fp t, u, ZZZ3;
    quad_mul4(q, Rr, fp_sub(Q, r.X), S1, PPP, ZZZ12, PPP, ZZZ12, PPP, t, u, ZZZ3, d0);
    r.Y = fp_sub(t, u);
    r.ZZ = ZZ3;
    r.ZZZ = ZZZ3;
    return r;
}
KZ_HD G1Xyzz xyzz_neg(const G1Xyzz& p) { return {p.X, fp_neg(p.Y), p.ZZ, p.ZZZ}; }

// [|x|]P for the BLS parameter, MSB first: 63 doublings and 5 additions of P
KZ_HD G1Xyzz quad_mul_xabs(Quad& q, const G1Xyzz& p) {
    G1Xyzz acc = p;
    const u64 k = ((u64)X_ABS_HI << 32) | X_ABS_LO;
    for (int i = 62; i >= 0; --i) {
        acc = quad_xyzz_dbl(q, acc);
        if ((k >> i) & 1) acc = quad_xyzz_add(q, acc, p);
    }
    return acc;
}
// T in G1  <=>  sigma(T) == -[x^2]T with sigma(x, y) = (beta x, y); projective comparison, no inversion:
//   beta X_T ZZ_Q == X_Q ZZ_T   and   Y_T ZZZ_Q == -Y_Q ZZZ_T.   Infinity is in G1.
KZ_HD bool quad_sum_in_g1(Quad& q, const G1Xyzz& t) {
    if (xyzz_is_inf(t)) return true;
    G1Xyzz Q = quad_mul_xabs(q, quad_mul_xabs(q, t));
    if (xyzz_is_inf(Q)) return false;
    Fp a, b, c, d;
    quad_mul4(q, fp_mul(fp_const(FP_BETA), t.X), Q.ZZ, Q.X, t.ZZ, t.Y, Q.ZZZ, Q.Y, t.ZZZ, a, b, c, d);
    return fp_eq(a, b) && fp_eq(c, fp_neg(d));
}
