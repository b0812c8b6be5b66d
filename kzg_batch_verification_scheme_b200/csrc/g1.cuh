// G1 group law for E: y^2 = x^3 + 4 over Fp -- affine, Jacobian (doubling chains) and XYZZ (bucket
// accumulation) -- plus the per-point hot path of BASELINE.json:5 item (b): decompression of the 48-byte
// ZCash encoding, on-curve check and subgroup check.  All exceptional cases (P+P, P-P, infinity) are
// handled explicitly so adversarial inputs (small-order points, duplicates) give exact results.
#pragma once
#include "field.cuh"

struct G1Aff { Fp x, y; };                  // infinity is encoded out of band (x = y = 0 never on curve)
struct G1Jac { Fp X, Y, Z; };               // Z == 0 <=> infinity
struct G1Xyzz { Fp X, Y, ZZ, ZZZ; };        // ZZ == 0 <=> infinity; x = X/ZZ, y = Y/ZZZ

KZ_HD bool aff_is_inf(const G1Aff& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
KZ_HD G1Aff aff_inf() { return {fp_zero(), fp_zero()}; }
KZ_HD G1Jac jac_inf() { return {fp_one(), fp_one(), fp_zero()}; }
KZ_HD bool jac_is_inf(const G1Jac& p) { return fp_is_zero(p.Z); }
KZ_HD G1Jac jac_from_aff(const G1Aff& p) {
    if (aff_is_inf(p)) return jac_inf();
    return {p.x, p.y, fp_one()};
}

// dbl-2009-l (a = 0): 2M + 5S.  Z = 0 maps to Z = 0; no point of order 2 exists on E(Fp).
KZ_HD G1Jac jac_dbl(const G1Jac& p) {
    Fp A = fp_sqr(p.X), B = fp_sqr(p.Y), C = fp_sqr(B);
    Fp t = fp_add(p.X, B);
    Fp D = fp_sub(fp_sub(fp_sqr(t), A), C);
    D = fp_dbl(D);
    Fp E = fp_add(fp_dbl(A), A);
    Fp F = fp_sqr(E);
    G1Jac r;
    r.X = fp_sub(F, fp_dbl(D));
    Fp C8 = fp_dbl(fp_dbl(fp_dbl(C)));
    r.Y = fp_sub(fp_mul(E, fp_sub(D, r.X)), C8);
    r.Z = fp_dbl(fp_mul(p.Y, p.Z));
    return r;
}
// Jacobian + affine (q not infinity): 8M + 3S
KZ_HD G1Jac jac_madd(const G1Jac& p, const G1Aff& q) {
    if (jac_is_inf(p)) return {q.x, q.y, fp_one()};
    Fp Z1Z1 = fp_sqr(p.Z);
    Fp U2 = fp_mul(q.x, Z1Z1);
    Fp S2 = fp_mul(fp_mul(q.y, p.Z), Z1Z1);
    Fp H = fp_sub(U2, p.X), Rr = fp_sub(S2, p.Y);
    if (fp_is_zero(H)) return fp_is_zero(Rr) ? jac_dbl(p) : jac_inf();
    Fp HH = fp_sqr(H), HHH = fp_mul(H, HH), V = fp_mul(p.X, HH);
    G1Jac r;
    r.X = fp_sub(fp_sub(fp_sqr(Rr), HHH), fp_dbl(V));
    r.Y = fp_sub(fp_mul(Rr, fp_sub(V, r.X)), fp_mul(p.Y, HHH));
    r.Z = fp_mul(p.Z, H);
    return r;
}
// Jacobian + Jacobian: 12M + 4S
KZ_HD G1Jac jac_add(const G1Jac& p, const G1Jac& q) {
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    Fp Z1Z1 = fp_sqr(p.Z), Z2Z2 = fp_sqr(q.Z);
    Fp U1 = fp_mul(p.X, Z2Z2), U2 = fp_mul(q.X, Z1Z1);
    Fp S1 = fp_mul(fp_mul(p.Y, q.Z), Z2Z2), S2 = fp_mul(fp_mul(q.Y, p.Z), Z1Z1);
    Fp H = fp_sub(U2, U1), Rr = fp_sub(S2, S1);
    if (fp_is_zero(H)) return fp_is_zero(Rr) ? jac_dbl(p) : jac_inf();
    Fp HH = fp_sqr(H), HHH = fp_mul(H, HH), V = fp_mul(U1, HH);
    G1Jac r;
    r.X = fp_sub(fp_sub(fp_sqr(Rr), HHH), fp_dbl(V));
    r.Y = fp_sub(fp_mul(Rr, fp_sub(V, r.X)), fp_mul(S1, HHH));
    r.Z = fp_mul(fp_mul(p.Z, q.Z), H);
    return r;
}
KZ_HD G1Jac jac_neg(const G1Jac& p) { return {p.X, fp_neg(p.Y), p.Z}; }
// the same test for a Jacobian point J (not infinity), without an inversion:
// beta X_J Z_Q^2 == X_Q Z_J^2  and  Y_J Z_Q^3 == -Y_Q Z_J^3
KZ_HD bool g1_subgroup_compare_jac(const G1Jac& j, const G1Jac& q) {
    if (jac_is_inf(q)) return false;
    Fp zq2 = fp_sqr(q.Z), zj2 = fp_sqr(j.Z);
    if (!fp_eq(fp_mul(fp_mul(fp_const(FP_BETA), j.X), zq2), fp_mul(q.X, zj2))) return false;
    return fp_eq(fp_mul(j.Y, fp_mul(zq2, q.Z)), fp_neg(fp_mul(q.Y, fp_mul(zj2, j.Z))));
}
KZ_COLD G1Aff jac_to_aff(const G1Jac& p) {        // one inversion; infinity -> (0,0)
    if (jac_is_inf(p)) return aff_inf();
    Fp zi = fp_inv(p.Z), zi2 = fp_sqr(zi);
    return {fp_mul(p.X, zi2), fp_mul(fp_mul(p.Y, zi2), zi)};
}
// [k]P, k = nlimbs x u32 little-endian scalar (thread-local), plain double-and-add (cold paths only)
KZ_COLD G1Jac jac_mul_limbs(const G1Jac& p, const u32* k, int nlimbs) {
    G1Jac r = jac_inf();
    for (int i = nlimbs * 32 - 1; i >= 0; --i) {
        r = jac_dbl(r);
        if ((k[i >> 5] >> (i & 31)) & 1) r = jac_add(r, p);
    }
    return r;
}

// ------------------------------------------------------------------ XYZZ
KZ_HD G1Xyzz xyzz_inf() { return {fp_zero(), fp_zero(), fp_zero(), fp_zero()}; }
KZ_HD bool xyzz_is_inf(const G1Xyzz& p) { return fp_is_zero(p.ZZ); }
// doubling of an affine point into XYZZ (mdbl-2008-s-1, a = 0)
KZ_HD G1Xyzz xyzz_dbl_aff(const G1Aff& q) {
    Fp U = fp_dbl(q.y), V = fp_sqr(U), W = fp_mul(U, V), S = fp_mul(q.x, V);
    Fp x2 = fp_sqr(q.x);
    Fp M = fp_add(fp_dbl(x2), x2);
    G1Xyzz r;
    r.X = fp_sub(fp_sqr(M), fp_dbl(S));
    r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, q.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}
KZ_HD G1Xyzz xyzz_dbl(const G1Xyzz& p) {         // dbl-2008-s-1
    if (xyzz_is_inf(p)) return p;
    Fp U = fp_dbl(p.Y), V = fp_sqr(U), W = fp_mul(U, V), S = fp_mul(p.X, V);
    Fp x2 = fp_sqr(p.X);
    Fp M = fp_add(fp_dbl(x2), x2);
    G1Xyzz r;
    r.X = fp_sub(fp_sqr(M), fp_dbl(S));
    r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.Y));
    r.ZZ = fp_mul(V, p.ZZ);
    r.ZZZ = fp_mul(W, p.ZZZ);
    return r;
}
// XYZZ + affine (q not infinity): madd-2008-s, 8M + 2S
KZ_HD G1Xyzz xyzz_madd(const G1Xyzz& p, const G1Aff& q) {
    if (xyzz_is_inf(p)) return {q.x, q.y, fp_one(), fp_one()};
    Fp U2 = fp_mul(q.x, p.ZZ), S2 = fp_mul(q.y, p.ZZZ);
    Fp Pp = fp_sub(U2, p.X), Rr = fp_sub(S2, p.Y);
    if (fp_is_zero(Pp)) return fp_is_zero(Rr) ? xyzz_dbl_aff(q) : xyzz_inf();
    Fp PP = fp_sqr(Pp), PPP = fp_mul(Pp, PP), Q = fp_mul(p.X, PP);
    G1Xyzz r;
    r.X = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Q));
    r.Y = fp_sub(fp_mul(Rr, fp_sub(Q, r.X)), fp_mul(p.Y, PPP));
    r.ZZ = fp_mul(p.ZZ, PP);
    r.ZZZ = fp_mul(p.ZZZ, PPP);
    return r;
}
// XYZZ + XYZZ: add-2008-s, 12M + 2S
KZ_HD G1Xyzz xyzz_add(const G1Xyzz& p, const G1Xyzz& q) {
    if (xyzz_is_inf(p)) return q;
    if (xyzz_is_inf(q)) return p;
    Fp U1 = fp_mul(p.X, q.ZZ), U2 = fp_mul(q.X, p.ZZ);
    Fp S1 = fp_mul(p.Y, q.ZZZ), S2 = fp_mul(q.Y, p.ZZZ);
    Fp Pp = fp_sub(U2, U1), Rr = fp_sub(S2, S1);
    if (fp_is_zero(Pp)) return fp_is_zero(Rr) ? xyzz_dbl(p) : xyzz_inf();
    Fp PP = fp_sqr(Pp), PPP = fp_mul(Pp, PP), Q = fp_mul(U1, PP);
    G1Xyzz r;
    r.X = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Q));
    r.Y = fp_sub(fp_mul(Rr, fp_sub(Q, r.X)), fp_mul(S1, PPP));
    r.ZZ = fp_mul(fp_mul(p.ZZ, q.ZZ), PP);
    r.ZZZ = fp_mul(fp_mul(p.ZZZ, q.ZZZ), PPP);
    return r;
}
// XYZZ -> Jacobian without inversion: (X*ZZ, Y*ZZZ, ZZ)   [x = X ZZ / ZZ^2, y = Y ZZZ / ZZ^3 as ZZ^3 = ZZZ^2]
KZ_HD G1Jac xyzz_to_jac(const G1Xyzz& p) {
    if (xyzz_is_inf(p)) return jac_inf();
    return {fp_mul(p.X, p.ZZ), fp_mul(p.Y, p.ZZZ), p.ZZ};
}
KZ_HD G1Xyzz xyzz_from_jac(const G1Jac& p) {
    if (jac_is_inf(p)) return xyzz_inf();
    Fp zz = fp_sqr(p.Z);
    return {p.X, p.Y, zz, fp_mul(zz, p.Z)};
}

// ------------------------------------------------------------------ serialization helpers
KZ_HD void aff_to_be96(u8* out, const G1Aff& p) {          // canonical x||y; infinity -> zeros
    fp_to_be(out, p.x);
    fp_to_be(out + 48, p.y);
}
KZ_HD bool aff_from_be96(G1Aff& p, const u8* in) { return fp_from_be(p.x, in) & fp_from_be(p.y, in + 48); }

// ------------------------------------------------------------------ sqrt:  a^((p+1)/4), sliding 4-bit windows
// 376 squarings + 78 table multiplications + 8 products to build the odd-power table (a, a^3, .., a^15).
// The schedule is a compile-time constant (same for every thread), so the table lives in local memory
// with uniform, coalesced indexing and all branches are warp-uniform.
// The chain is products only, so it runs on lazy values in [0, 2p) (field.cuh) with one reduction at the end.
KZ_HD Fp fp_sqrt_candidate(const Fp& a) {
    Fp tab[8];
    tab[0] = a;
    Fp a2 = fp_sqr_lazy(a);
    for (int i = 1; i < 8; ++i) tab[i] = fp_mul_lazy(tab[i - 1], a2);
    Fp r = tab[SQRT_SCHED[1]];                       // first step: leading window, no squarings
    for (int s = 1; s < SQRT_SCHED_STEPS; ++s) {
        int nsq = SQRT_SCHED[2 * s], idx = SQRT_SCHED[2 * s + 1];
        for (int k = 0; k < nsq; ++k) r = fp_sqr_lazy(r);
        if (idx != 0xFF) r = fp_mul_lazy(r, tab[idx]);
    }
    return fp_reduce_once(r);
}

// ------------------------------------------------------------------ subgroup check
// [|x|]P for the BLS parameter |x| = 0xd201000000010000 (64 bits, Hamming weight 6), MSB-first.
KZ_HD G1Jac jac_mul_xabs_aff(const G1Aff& p) {
    G1Jac acc = {p.x, p.y, fp_one()};
    const u64 k = ((u64)X_ABS_HI << 32) | X_ABS_LO;
    for (int i = 62; i >= 0; --i) {
        acc = jac_dbl(acc);
        if ((k >> i) & 1) acc = jac_madd(acc, p);
    }
    return acc;
}
KZ_HD G1Jac jac_mul_xabs(const G1Jac& p) {
    G1Jac acc = p;
    const u64 k = ((u64)X_ABS_HI << 32) | X_ABS_LO;
    for (int i = 62; i >= 0; --i) {
        acc = jac_dbl(acc);
        if ((k >> i) & 1) acc = jac_add(acc, p);
    }
    return acc;
}
// P in G1  <=>  sigma(P) == -[x^2]P with sigma(x,y) = (beta x, y)   (P != infinity)
KZ_HD bool g1_subgroup_compare(const G1Aff& p, const G1Jac& q);
KZ_HD bool g1_in_subgroup(const G1Aff& p) { return g1_subgroup_compare(p, jac_mul_xabs(jac_mul_xabs_aff(p))); }

// ------------------------------------------------------------------ decompress + validate (K1 body)
enum { ST_OK = 0, ST_BAD_FLAGS = 1, ST_X_GE_P = 2, ST_NOT_ON_CURVE = 3, ST_NOT_IN_G1 = 4 };

// in: 48 bytes big-endian as 12 big-endian-loaded words w[0..11] (w[0] holds bytes 0..3).
// Stage 1 (flags, range, square root, sign): out = affine point in Montgomery form, (0,0) for infinity or
// any failure.  Returns the status byte; ST_OK here still lacks the subgroup check.
KZ_HD u32 g1_decompress_sqrt(G1Aff& out, const u32* w) {
    out = aff_inf();
    u32 b0 = w[0] >> 24;
    if (!(b0 & 0x80)) return ST_BAD_FLAGS;
    if (b0 & 0x40) {
        u32 rest = w[0] & 0x3FFFFFFFu;
        KZ_UNROLL for (int i = 1; i < 12; ++i) rest |= w[i];
        return rest ? ST_BAD_FLAGS : ST_OK;
    }
    Fp raw;
    KZ_UNROLL for (int i = 0; i < 12; ++i) raw.v[i] = w[11 - i];
    raw.v[11] &= 0x1FFFFFFFu;
    u32 pl[12];
    KZ_UNROLL for (int i = 0; i < 12; ++i) pl[i] = FP_P[i];
    if (limbs_ge12(raw.v, pl)) return ST_X_GE_P;
    Fp x = fp_to_mont(raw);
    Fp rhs = fp_add(fp_mul(fp_sqr(x), x), fp_const(FP_B));
    Fp y = fp_sqrt_candidate(rhs);
    if (!fp_eq(fp_sqr(y), rhs)) return ST_NOT_ON_CURVE;
    if (fp_is_lex_largest(y) != ((b0 & 0x20) != 0)) y = fp_neg(y);
    out = {x, y};
    return ST_OK;
}
// Stage 3 of the subgroup check: given Q = [x^2]P (Jacobian), decide sigma(P) == -Q.
KZ_HD bool g1_subgroup_compare(const G1Aff& p, const G1Jac& q) {
    if (jac_is_inf(q)) return false;
    Fp zz = fp_sqr(q.Z);
    Fp bx = fp_mul(fp_const(FP_BETA), p.x);
    if (!fp_eq(q.X, fp_mul(bx, zz))) return false;
    Fp zzz = fp_mul(zz, q.Z);
    return fp_eq(q.Y, fp_neg(fp_mul(p.y, zzz)));
}
// fused form (debug operator, setup point, emulation)
KZ_HD u32 g1_decompress_validate(G1Aff& out, const u32* w, bool check_subgroup = true) {
    G1Aff p;
    u32 st = g1_decompress_sqrt(p, w);
    out = aff_inf();
    if (st != ST_OK) return st;
    if (aff_is_inf(p)) return ST_OK;
    if (check_subgroup && !g1_in_subgroup(p)) return ST_NOT_IN_G1;
    out = p;
    return ST_OK;
}
// compress (cold path: synthetic generator)
KZ_HD void g1_compress_words(u32* w, const G1Aff& p) {     // w[0..11] big-endian words
    if (aff_is_inf(p)) {
        w[0] = 0xC0000000u;
        KZ_UNROLL for (int i = 1; i < 12; ++i) w[i] = 0;
        return;
    }
    Fp c = fp_from_mont(p.x);
    KZ_UNROLL for (int i = 0; i < 12; ++i) w[i] = c.v[11 - i];
    w[0] |= 0x80000000u;
    if (fp_is_lex_largest(p.y)) w[0] |= 0x20000000u;
}
