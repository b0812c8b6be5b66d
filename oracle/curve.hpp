// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  G1 / G2 group arithmetic (Jacobian), ZCash-format
// (de)compression, subgroup checks, and the two-pairing product check.
// Semantics: BASELINE.json:5 + SURVEY.md App. A/B (upstream reference is LICENSE-only).
#pragma once
#include <vector>

#include "tower.hpp"

namespace orc {

constexpr u64 X_ABS = 0xd201000000010000ull;   // BLS parameter x = -X_ABS

// ------------------------------------------------------------------ generic short-Weierstrass a=0
template <class F>
struct Jac {
    F X, Y, Z;
    static Jac inf() { return {F::one(), F::one(), F::zero()}; }
    static Jac from_affine(const F& x, const F& y) { return {x, y, F::one()}; }
    bool is_inf() const { return Z.is_zero(); }
    Jac neg() const { return {X, -Y, Z}; }
    Jac dbl() const {
        if (is_inf()) return *this;
        F A = X.sqr(), B = Y.sqr(), C = B.sqr();
        F D = ((X + B).sqr() - A - C).dbl();
        F E = A.dbl() + A;
        F Fv = E.sqr();
        F X3 = Fv - D.dbl();
        F Y3 = E * (D - X3) - C.dbl().dbl().dbl();
        F Z3 = (Y * Z).dbl();
        return {X3, Y3, Z3};
    }
    Jac add(const Jac& o) const {
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        F U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        F S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        F H = U2 - U1, Rr = S2 - S1;
        if (H.is_zero()) return Rr.is_zero() ? dbl() : inf();
        F HH = H.sqr(), HHH = H * HH, V = U1 * HH;
        F X3 = Rr.sqr() - HHH - V.dbl();
        F Y3 = Rr * (V - X3) - S1 * HHH;
        F Z3 = Z * o.Z * H;
        return {X3, Y3, Z3};
    }
    // scalar as little-endian u64 limbs
    Jac mul(const u64* k, int nl) const {
        Jac r = inf();
        for (int i = nl * 64 - 1; i >= 0; --i) {
            r = r.dbl();
            if (k[i / 64] >> (i % 64) & 1) r = r.add(*this);
        }
        return r;
    }
    bool to_affine(F& x, F& y) const {     // false for infinity
        if (is_inf()) return false;
        F zi = Z.inv(), zi2 = zi.sqr();
        x = X * zi2;
        y = Y * zi2 * zi;
        return true;
    }
    bool eq(const Jac& o) const {
        if (is_inf() || o.is_inf()) return is_inf() && o.is_inf();
        F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        return X * Z2Z2 == o.X * Z1Z1 && Y * o.Z * Z2Z2 == o.Y * Z * Z1Z1;
    }
};
using G1J = Jac<Fp>;
using G2J = Jac<Fp2>;

struct G1A {           // affine; inf flag explicit
    Fp x, y;
    bool inf;
    static G1A infinity() { return {Fp::zero(), Fp::zero(), true}; }
    G1J jac() const { return inf ? G1J::inf() : G1J::from_affine(x, y); }
};
inline G1A g1_affine(const G1J& p) {
    G1A a = G1A::infinity();
    if (p.to_affine(a.x, a.y)) a.inf = false;
    return a;
}
// Montgomery batch normalisation
inline void g1_batch_affine(const G1J* in, G1A* out, size_t n) {
    std::vector<Fp> pre(n);
    Fp acc = Fp::one();
    for (size_t i = 0; i < n; ++i) {
        pre[i] = acc;
        if (!in[i].is_inf()) acc = acc * in[i].Z;
    }
    Fp inv = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (in[i].is_inf()) { out[i] = G1A::infinity(); continue; }
        Fp zi = inv * pre[i];
        inv = inv * in[i].Z;
        Fp zi2 = zi.sqr();
        out[i] = {in[i].X * zi2, in[i].Y * zi2 * zi, false};
    }
}

inline Fp fp_from_hex(const char* hex) {
    u64 raw[6] = {0};
    int len = (int)strlen(hex);
    for (int i = 0; i < len; ++i) {
        char ch = hex[len - 1 - i];
        u64 v = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
        raw[i / 16] |= v << (4 * (i % 16));
    }
    return Fp::from_raw(raw);
}
inline const G1A& g1_generator() {
    static const G1A g = {
        fp_from_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"),
        fp_from_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"),
        false};
    return g;
}
struct G2A {
    Fp2 x, y;
    bool inf;
};
inline const G2A& g2_generator() {
    static const G2A g = {
        {fp_from_hex("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"),
         fp_from_hex("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e")},
        {fp_from_hex("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801"),
         fp_from_hex("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be")},
        false};
    return g;
}
inline Fp fp_b() { return Fp::from_u64(4); }
inline Fp2 fp2_b_twist() { return Fp2{Fp::from_u64(4), Fp::from_u64(4)}; }   // 4(1+u)

// ------------------------------------------------------------------ G1 subgroup membership
// beta = 2^((p-1)/3): primitive cube root of unity, computed, not transcribed.
inline const Fp& g1_beta() {
    static const Fp beta = [] {
        u64 e[6], onev[6] = {1};
        sub_raw<6>(e, fp_params().mod, onev);
        u128 rem = 0;
        for (int i = 5; i >= 0; --i) {
            u128 cur = (rem << 64) | e[i];
            e[i] = (u64)(cur / 3);
            rem = cur % 3;
        }
        return Fp::from_u64(2).pow(e, 6);
    }();
    return beta;
}
// fast test: sigma(P) == -[x^2]P, sigma(x,y) = (beta x, y)
inline bool g1_in_subgroup_fast(const G1A& p) {
    if (p.inf) return true;
    u64 k[1] = {X_ABS};
    G1J q = p.jac().mul(k, 1).mul(k, 1);          // [x^2]P
    G1J s = G1J::from_affine(g1_beta() * p.x, -p.y);   // -sigma(P)
    return q.eq(s);
}
// slow test: [r]P == O
inline bool g1_in_subgroup_slow(const G1A& p) {
    if (p.inf) return true;
    return p.jac().mul(fr_params().mod, 4).is_inf();
}

// ------------------------------------------------------------------ serialization
enum : u8 { ST_OK = 0, ST_BAD_FLAGS = 1, ST_X_GE_P = 2, ST_NOT_ON_CURVE = 3, ST_NOT_IN_G1 = 4 };

inline void g1_compress(u8 out[48], const G1A& p) {
    if (p.inf) { memset(out, 0, 48); out[0] = 0xC0; return; }
    p.x.to_bytes_be(out);
    out[0] |= 0x80;
    if (p.y.is_lex_largest()) out[0] |= 0x20;
}
// Validation precedence per SURVEY App. B.2.  `check_subgroup`: 0 none, 1 fast, 2 slow.
inline u8 g1_decompress(G1A& out, const u8 in[48], int check_subgroup = 1) {
    out = G1A::infinity();
    u8 b0 = in[0];
    if (!(b0 & 0x80)) return ST_BAD_FLAGS;
    if (b0 & 0x40) {
        if (b0 != 0xC0) return ST_BAD_FLAGS;
        for (int i = 1; i < 48; ++i) if (in[i]) return ST_BAD_FLAGS;
        return ST_OK;
    }
    u8 tmp[48];
    memcpy(tmp, in, 48);
    tmp[0] &= 0x1F;
    Fp x;
    if (!Fp::from_bytes_be(x, tmp)) return ST_X_GE_P;
    Fp rhs = x.sqr() * x + fp_b();
    // y = rhs^((p+1)/4)
    u64 e[6], onev[6] = {1};
    add_raw<6>(e, fp_params().mod, onev);
    for (int i = 0; i < 6; ++i) e[i] = (e[i] >> 2) | (i < 5 ? e[i + 1] << 62 : 0);
    Fp y = rhs.pow(e, 6);
    if (y.sqr() != rhs) return ST_NOT_ON_CURVE;
    if (y.is_lex_largest() != bool(b0 & 0x20)) y = -y;
    G1A p{x, y, false};
    if (check_subgroup == 1 && !g1_in_subgroup_fast(p)) return ST_NOT_IN_G1;
    if (check_subgroup == 2 && !g1_in_subgroup_slow(p)) return ST_NOT_IN_G1;
    out = p;
    return ST_OK;
}
// canonical 96-byte x||y, infinity = zeros
inline void g1_affine_bytes(u8 out[96], const G1A& p) {
    if (p.inf) { memset(out, 0, 96); return; }
    p.x.to_bytes_be(out);
    p.y.to_bytes_be(out + 48);
}
inline bool g1_from_affine_bytes(G1A& out, const u8 in[96]) {
    bool allz = true;
    for (int i = 0; i < 96; ++i) allz &= in[i] == 0;
    if (allz) { out = G1A::infinity(); return true; }
    out.inf = false;
    return Fp::from_bytes_be(out.x, in) && Fp::from_bytes_be(out.y, in + 48);
}

inline void g2_compress(u8 out[96], const G2A& p) {
    if (p.inf) { memset(out, 0, 96); out[0] = 0xC0; return; }
    p.x.c1.to_bytes_be(out);
    p.x.c0.to_bytes_be(out + 48);
    out[0] |= 0x80;
    if (p.y.is_lex_largest()) out[0] |= 0x20;
}
// returns false on any malformed / off-curve / off-subgroup input (setup points are all-or-nothing)
inline bool g2_decompress(G2A& out, const u8 in[96]) {
    u8 b0 = in[0];
    if (!(b0 & 0x80) || (b0 & 0x40)) return false;
    u8 tmp[48];
    memcpy(tmp, in, 48);
    tmp[0] &= 0x1F;
    Fp2 x;
    if (!Fp::from_bytes_be(x.c1, tmp) || !Fp::from_bytes_be(x.c0, in + 48)) return false;
    Fp2 rhs = x.sqr() * x + fp2_b_twist();
    Fp2 y;
    if (!rhs.sqrt(y)) return false;
    if (y.is_lex_largest() != bool(b0 & 0x20)) y = -y;
    out = {x, y, false};
    return G2J::from_affine(x, y).mul(fr_params().mod, 4).is_inf();
}

// ------------------------------------------------------------------ pairing
// Line through twist points evaluated at P=(xP,yP) in E(Fp), scaled by w^3 (subfield factor):
//   l = (lam*xT - yT) + (-lam*xP) v + yP v w     (derivation in DESIGN.md "Pairing")
inline Fp12 line_value(const Fp2& lam, const Fp2& xT, const Fp2& yT, const Fp& xP, const Fp& yP) {
    Fp12 l{Fp6::zero(), Fp6::zero()};
    l.c0.c0 = lam * xT - yT;
    l.c0.c1 = -(lam.mul_fp(xP));
    l.c1.c1 = Fp2{yP, Fp::zero()};
    return l;
}
// product of Miller functions f_{|x|,Q_k}(P_k), conjugated for x<0.  Affine twist arithmetic.
inline Fp12 miller_loop_multi(const G1A* P, const G2A* Q, int n) {
    Fp12 f = Fp12::one();
    std::vector<Fp2> tx(n), ty(n);
    for (int k = 0; k < n; ++k) { tx[k] = Q[k].x; ty[k] = Q[k].y; }
    for (int i = 62; i >= 0; --i) {
        f = f.sqr();
        for (int k = 0; k < n; ++k) {
            if (P[k].inf || Q[k].inf) continue;
            Fp2 x2 = tx[k].sqr();
            Fp2 lam = (x2.dbl() + x2) * ty[k].dbl().inv();
            f = f * line_value(lam, tx[k], ty[k], P[k].x, P[k].y);
            Fp2 x3 = lam.sqr() - tx[k].dbl();
            ty[k] = lam * (tx[k] - x3) - ty[k];
            tx[k] = x3;
        }
        if (X_ABS >> i & 1) {
            for (int k = 0; k < n; ++k) {
                if (P[k].inf || Q[k].inf) continue;
                Fp2 lam = (ty[k] - Q[k].y) * (tx[k] - Q[k].x).inv();
                f = f * line_value(lam, Q[k].x, Q[k].y, P[k].x, P[k].y);
                Fp2 x3 = lam.sqr() - tx[k] - Q[k].x;
                ty[k] = lam * (tx[k] - x3) - ty[k];
                tx[k] = x3;
            }
        }
    }
    return f.conj();
}
inline Fp12 pow_x_abs(const Fp12& a) {
    u64 k[1] = {X_ABS};
    return a.pow(k, 1);
}
// f^(3 (p^12-1)/r):  easy part, then hard part 3(p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3
// (identity checked numerically in tests/test_pymodel.py).  x<0: f^x = conj(f^|x|) in the
// cyclotomic subgroup.
inline Fp12 final_exp(const Fp12& f0) {
    Fp12 f = f0.conj() * f0.inv();
    f = frob2(f) * f;
    auto powx = [](const Fp12& a) { return pow_x_abs(a).conj(); };       // a^x
    Fp12 a = powx(f) * f.conj();            // f^(x-1)
    a = powx(a) * a.conj();                 // ^(x-1)
    Fp12 b = powx(a) * frob1(a);            // ^(x+p)
    Fp12 c = powx(powx(b)) * frob2(b) * b.conj();   // ^(x^2+p^2-1)
    return c * f.sqr() * f;
}
inline bool pairing_product_is_one(const G1A* P, const G2A* Q, int n) {
    return final_exp(miller_loop_multi(P, Q, n)) == Fp12::one();
}

}  // namespace orc
