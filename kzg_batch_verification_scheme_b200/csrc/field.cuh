// Fp (381-bit, 12 x u32) and Fr (255-bit, 8 x u32) Montgomery arithmetic for sm_100a.
// BASELINE.json:5 item (a).  All values are kept canonical (< p) in Montgomery form R = 2^384 / 2^256.
// The multiply/add/sub carry chains are the generated inline-PTX blocks of mont_gen.cuh.
#pragma once
#include "common.cuh"
#include "constants_gen.cuh"
#if !defined(KZGB_EMU)
#include "mont_gen.cuh"
#endif

struct alignas(16) Fp { u32 v[12]; };     // 16-byte aligned: 128-bit loads/stores for global, shared and local copies
struct alignas(16) Fr { u32 v[8]; };

#if defined(KZGB_EMU)
// ---- portable limb loops (host emulation of the PTX blocks; tests only)
template <int N>
inline void emu_mont_mul(u32* r, const u32* a, const u32* b, const u32* p, u32 m0) {
    u32 t[N + 2] = {0};
    for (int i = 0; i < N; ++i) {
        u64 c = 0;
        for (int j = 0; j < N; ++j) { u64 v = (u64)a[j] * b[i] + t[j] + c; t[j] = (u32)v; c = v >> 32; }
        u64 v = (u64)t[N] + c; t[N] = (u32)v; t[N + 1] = (u32)(v >> 32);
        u32 q = t[0] * m0;
        v = (u64)q * p[0] + t[0]; c = v >> 32;
        for (int j = 1; j < N; ++j) { v = (u64)q * p[j] + t[j] + c; t[j - 1] = (u32)v; c = v >> 32; }
        v = (u64)t[N] + c; t[N - 1] = (u32)v; t[N] = t[N + 1] + (u32)(v >> 32);
    }
    u32 s[N]; u64 bw = 0;
    for (int j = 0; j < N; ++j) { u64 v = (u64)t[j] - p[j] - bw; s[j] = (u32)v; bw = (v >> 32) & 1; }
    bool ge = t[N] || !bw;
    for (int j = 0; j < N; ++j) r[j] = ge ? s[j] : t[j];
}
template <int N>
inline void emu_add(u32* r, const u32* a, const u32* b, const u32* p) {
    u32 t[N], s[N]; u64 c = 0, bw = 0;
    for (int j = 0; j < N; ++j) { u64 v = (u64)a[j] + b[j] + c; t[j] = (u32)v; c = v >> 32; }
    for (int j = 0; j < N; ++j) { u64 v = (u64)t[j] - p[j] - bw; s[j] = (u32)v; bw = (v >> 32) & 1; }
    for (int j = 0; j < N; ++j) r[j] = bw ? t[j] : s[j];
}
template <int N>
inline void emu_sub(u32* r, const u32* a, const u32* b, const u32* p) {
    u32 t[N]; u64 bw = 0, c = 0;
    for (int j = 0; j < N; ++j) { u64 v = (u64)a[j] - b[j] - bw; t[j] = (u32)v; bw = (v >> 32) & 1; }
    for (int j = 0; j < N; ++j) { u64 v = (u64)t[j] + (bw ? p[j] : 0) + c; r[j] = (u32)v; c = v >> 32; }
}
#endif

// ------------------------------------------------------------------ Fp
KZ_HD Fp fp_mul(const Fp& a, const Fp& b) {
    Fp r;
#if defined(KZGB_EMU)
    emu_mont_mul<12>(r.v, a.v, b.v, FP_P, FP_M0);
#else
    fp_mont_mul_ptx(r.v, a.v, b.v);
    fp_reduce_ptx(r.v);
#endif
    return r;
}
KZ_HD Fp fp_sqr(const Fp& a) {
#if defined(KZGB_EMU)
    return fp_mul(a, a);
#else
    Fp r;
    fp_mont_sqr_ptx(r.v, a.v);          // dedicated squaring: 234 instead of 300 wide multiply-adds
    fp_reduce_ptx(r.v);
    return r;
#endif
}
// Lazy variants for pure product chains (square root): operands and results in [0, 2p).  R = 2^384 > 4p, so a
// Montgomery product of operands below 2p is below 1.5p and the conditional subtraction can wait until the end
// of the chain (tools/gen_mont.py --selftest covers operands up to 2p - 1).  fp_reduce_once: [0, 2p) -> [0, p).
KZ_HD Fp fp_mul_lazy(const Fp& a, const Fp& b) {
#if defined(KZGB_EMU)
    return fp_mul(a, b);
#else
    Fp r;
    fp_mont_mul_ptx(r.v, a.v, b.v);
    return r;
#endif
}
KZ_HD Fp fp_sqr_lazy(const Fp& a) {
#if defined(KZGB_EMU)
    return fp_mul(a, a);
#else
    Fp r;
    fp_mont_sqr_ptx(r.v, a.v);
    return r;
#endif
}
KZ_HD Fp fp_reduce_once(Fp r) {
#if !defined(KZGB_EMU)
    fp_reduce_ptx(r.v);
#endif
    return r;
}
KZ_HD Fp fp_add(const Fp& a, const Fp& b) {
    Fp r;
#if defined(KZGB_EMU)
    emu_add<12>(r.v, a.v, b.v, FP_P);
#else
    fp_add_ptx(r.v, a.v, b.v);
#endif
    return r;
}
KZ_HD Fp fp_sub(const Fp& a, const Fp& b) {
    Fp r;
#if defined(KZGB_EMU)
    emu_sub<12>(r.v, a.v, b.v, FP_P);
#else
    fp_sub_ptx(r.v, a.v, b.v);
#endif
    return r;
}
KZ_HD Fp fp_zero() { Fp r; KZ_UNROLL for (int i = 0; i < 12; ++i) r.v[i] = 0; return r; }
KZ_HD Fp fp_const(const u32* c) { Fp r; KZ_UNROLL for (int i = 0; i < 12; ++i) r.v[i] = c[i]; return r; }
KZ_HD Fp fp_one() { return fp_const(FP_ONE); }
KZ_HD Fp fp_neg(const Fp& a) { return fp_sub(fp_zero(), a); }
KZ_HD Fp fp_dbl(const Fp& a) { return fp_add(a, a); }
KZ_HD bool fp_is_zero(const Fp& a) { u32 o = 0; KZ_UNROLL for (int i = 0; i < 12; ++i) o |= a.v[i]; return o == 0; }
KZ_HD bool fp_eq(const Fp& a, const Fp& b) { u32 o = 0; KZ_UNROLL for (int i = 0; i < 12; ++i) o |= a.v[i] ^ b.v[i]; return o == 0; }
// raw canonical limbs <-> Montgomery
KZ_HD Fp fp_to_mont(const Fp& raw) { return fp_mul(raw, fp_const(FP_R2)); }
KZ_HD Fp fp_from_mont(const Fp& a) { Fp o = fp_zero(); o.v[0] = 1; return fp_mul(a, o); }
// raw limb comparison a >= b
KZ_HD bool limbs_ge12(const u32* a, const u32* b) {
    u32 bw = 0;
    KZ_UNROLL for (int i = 0; i < 12; ++i) { u64 t = (u64)a[i] - b[i] - bw; bw = (u32)(t >> 32) & 1; }
    return bw == 0;
}
// canonical(a) > (p-1)/2
KZ_HD bool fp_is_lex_largest(const Fp& a) {
    Fp c = fp_from_mont(a);
    u32 half[12];
    KZ_UNROLL for (int i = 0; i < 12; ++i) half[i] = FP_HALF[i];
    return !limbs_ge12(half, c.v);
}
// a^e, e = nlimbs x u32 little-endian in constant memory (uniform control flow)
KZ_HD Fp fp_pow_const(const Fp& a, const u32* e, int nbits) {
    Fp r = a;                                   // top bit of e is bit nbits-1 and is set
    for (int i = nbits - 2; i >= 0; --i) {
        r = fp_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
    }
    return r;
}
// Inversion by the binary extended Euclidean algorithm on the canonical limbs (variable time: nothing here
// is secret).  ~760 shift/subtract steps instead of the ~570 Montgomery products of a Fermat power -- the
// inversion sits on latency-critical single-thread paths (pairing, Jacobian -> affine).  inv(0) = 0.
// Input a*R (Montgomery); the Euclid result (a R)^-1 is brought back with one product by R^3.
KZ_COLD Fp fp_inv(const Fp& a) {
    u32 orv = 0;
    for (int i = 0; i < 12; ++i) orv |= a.v[i];
    if (!orv) return fp_zero();
    u32 u[12], v[12], x1[12], x2[12], p[12];
    for (int i = 0; i < 12; ++i) { u[i] = a.v[i]; p[i] = FP_P[i]; v[i] = p[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    auto is_one = [](const u32* w) { u32 o = w[0] ^ 1u; for (int i = 1; i < 12; ++i) o |= w[i]; return o == 0; };
    auto shr1 = [](u32* w) { for (int i = 0; i < 11; ++i) w[i] = (w[i] >> 1) | (w[i + 1] << 31); w[11] >>= 1; };
    auto half_mod = [&](u32* w) {              // w <- w / 2 mod p
        if (w[0] & 1u) { u64 c = 0; for (int i = 0; i < 12; ++i) { u64 s = (u64)w[i] + p[i] + c; w[i] = (u32)s; c = s >> 32; } }
        for (int i = 0; i < 11; ++i) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
        w[11] >>= 1;
    };
    auto sub_to = [](u32* w, const u32* y) {   // w <- w - y, returns borrow
        u64 bw = 0;
        for (int i = 0; i < 12; ++i) { u64 s = (u64)w[i] - y[i] - bw; w[i] = (u32)s; bw = (s >> 32) & 1; }
        return (u32)bw;
    };
    auto sub_mod = [&](u32* w, const u32* y) { // w <- w - y mod p
        if (sub_to(w, y)) { u64 c = 0; for (int i = 0; i < 12; ++i) { u64 s = (u64)w[i] + p[i] + c; w[i] = (u32)s; c = s >> 32; } }
    };
    for (int guard = 0; guard < 2000 && !is_one(u) && !is_one(v); ++guard) {
        while (!(u[0] & 1u)) { shr1(u); half_mod(x1); }
        while (!(v[0] & 1u)) { shr1(v); half_mod(x2); }
        if (limbs_ge12(u, v)) { sub_to(u, v); sub_mod(x1, x2); }
        else { sub_to(v, u); sub_mod(x2, x1); }
    }
    Fp r;
    const u32* res = is_one(u) ? x1 : x2;
    for (int i = 0; i < 12; ++i) r.v[i] = res[i];
    return fp_mul(r, fp_const(FP_R3));
}
// big-endian 48 bytes <-> raw limbs
KZ_HD void fp_raw_from_be(Fp& r, const u8* b) {
    KZ_UNROLL for (int i = 0; i < 12; ++i) {
        const u8* q = b + 4 * (11 - i);
        r.v[i] = (u32)q[0] << 24 | (u32)q[1] << 16 | (u32)q[2] << 8 | q[3];
    }
}
KZ_HD void fp_raw_to_be(u8* b, const Fp& r) {
    KZ_UNROLL for (int i = 0; i < 12; ++i) {
        u8* q = b + 4 * (11 - i);
        q[0] = (u8)(r.v[i] >> 24); q[1] = (u8)(r.v[i] >> 16); q[2] = (u8)(r.v[i] >> 8); q[3] = (u8)r.v[i];
    }
}
KZ_HD bool fp_from_be(Fp& out, const u8* b) {      // false if >= p
    Fp raw; fp_raw_from_be(raw, b);
    u32 p[12];
    KZ_UNROLL for (int i = 0; i < 12; ++i) p[i] = FP_P[i];
    if (limbs_ge12(raw.v, p)) { out = fp_zero(); return false; }
    out = fp_to_mont(raw);
    return true;
}
KZ_HD void fp_to_be(u8* b, const Fp& a) { fp_raw_to_be(b, fp_from_mont(a)); }

// ------------------------------------------------------------------ Fr
KZ_HD Fr fr_mul(const Fr& a, const Fr& b) {
    Fr r;
#if defined(KZGB_EMU)
    emu_mont_mul<8>(r.v, a.v, b.v, FR_P, FR_M0);
#else
    fr_mont_mul_ptx(r.v, a.v, b.v);
    fr_reduce_ptx(r.v);
#endif
    return r;
}
KZ_HD Fr fr_add(const Fr& a, const Fr& b) {
    Fr r;
#if defined(KZGB_EMU)
    emu_add<8>(r.v, a.v, b.v, FR_P);
#else
    fr_add_ptx(r.v, a.v, b.v);
#endif
    return r;
}
KZ_HD Fr fr_sub(const Fr& a, const Fr& b) {
    Fr r;
#if defined(KZGB_EMU)
    emu_sub<8>(r.v, a.v, b.v, FR_P);
#else
    fr_sub_ptx(r.v, a.v, b.v);
#endif
    return r;
}
KZ_HD Fr fr_zero() { Fr r; KZ_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = 0; return r; }
KZ_HD Fr fr_const(const u32* c) { Fr r; KZ_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = c[i]; return r; }
KZ_HD Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }
KZ_HD bool fr_is_zero(const Fr& a) { u32 o = 0; KZ_UNROLL for (int i = 0; i < 8; ++i) o |= a.v[i]; return o == 0; }
KZ_HD Fr fr_to_mont(const Fr& raw) { return fr_mul(raw, fr_const(FR_R2)); }      // raw MUST be < r (the PTX product keeps no top carry)
// any 256-bit value -> [0, r): 2^256 < 3r, so two conditional subtractions suffice
KZ_HD Fr fr_reduce_raw(const Fr& a) {
    Fr r = a;
#if defined(KZGB_EMU)
    for (int rep = 0; rep < 2; ++rep) {
        u32 t[8]; u64 bw = 0;
        for (int j = 0; j < 8; ++j) { u64 v = (u64)r.v[j] - FR_P[j] - bw; t[j] = (u32)v; bw = (v >> 32) & 1; }
        if (!bw) for (int j = 0; j < 8; ++j) r.v[j] = t[j];
    }
#else
    fr_reduce_ptx(r.v);
    fr_reduce_ptx(r.v);
#endif
    return r;
}
KZ_HD Fr fr_from_mont(const Fr& a) { Fr o = fr_zero(); o.v[0] = 1; return fr_mul(a, o); }
KZ_HD bool limbs_ge8(const u32* a, const u32* b) {
    u32 bw = 0;
    KZ_UNROLL for (int i = 0; i < 8; ++i) { u64 t = (u64)a[i] - b[i] - bw; bw = (u32)(t >> 32) & 1; }
    return bw == 0;
}
KZ_HD void fr_raw_from_be(Fr& r, const u8* b) {
    KZ_UNROLL for (int i = 0; i < 8; ++i) {
        const u8* q = b + 4 * (7 - i);
        r.v[i] = (u32)q[0] << 24 | (u32)q[1] << 16 | (u32)q[2] << 8 | q[3];
    }
}
KZ_HD void fr_raw_to_be(u8* b, const Fr& r) {
    KZ_UNROLL for (int i = 0; i < 8; ++i) {
        u8* q = b + 4 * (7 - i);
        q[0] = (u8)(r.v[i] >> 24); q[1] = (u8)(r.v[i] >> 16); q[2] = (u8)(r.v[i] >> 8); q[3] = (u8)r.v[i];
    }
}
KZ_HD bool fr_raw_is_canonical(const Fr& raw) {
    u32 p[8];
    KZ_UNROLL for (int i = 0; i < 8; ++i) p[i] = FR_P[i];
    return !limbs_ge8(raw.v, p);
}
KZ_HD Fr fr_pow_const(const Fr& a, const u32* e, int nbits) {
    Fr r = a;
    for (int i = nbits - 2; i >= 0; --i) {
        r = fr_mul(r, r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = fr_mul(r, a);
    }
    return r;
}
KZ_COLD Fr fr_inv(const Fr& a) { return fr_pow_const(a, EXP_RM2, 255); }
