// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Extension tower Fp2 / Fp6 / Fp12 over BLS12-381 Fp.
//   Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - (1+u)), Fp12 = Fp6[w]/(w^2 - v)   
// Frobenius constants are computed at start-up by exponentiation (never transcribed).
// Upstream reference is LICENSE-only (/root/reference/LICENSE:1-201): nothing to follow there.
#pragma once
#include "field.hpp"

namespace orc {

struct Fp2 {
    Fp c0, c1;
    static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fp2& o) const { return !(*this == o); }
    Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fp2 operator-() const { return {-c0, -c1}; }
    Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fp2 operator*(const Fp2& o) const {
        Fp a = c0 * o.c0, b = c1 * o.c1;
        Fp c = (c0 + c1) * (o.c0 + o.c1);
        return {a - b, c - a - b};
    }
    Fp2 sqr() const {
        Fp a = (c0 + c1) * (c0 - c1);
        Fp b = c0 * c1;
        return {a, b.dbl()};
    }
    Fp2 mul_fp(const Fp& s) const { return {c0 * s, c1 * s}; }
    Fp2 conj() const { return {c0, -c1}; }
    Fp2 mul_xi() const { return {c0 - c1, c0 + c1}; }   // * (1+u)
    Fp2 inv() const {
        Fp n = (c0.sqr() + c1.sqr()).inv();
        return {c0 * n, -(c1 * n)};
    }
    Fp2 pow(const u64* e, int nl) const {
        Fp2 r = one();
        for (int i = nl * 64 - 1; i >= 0; --i) {
            r = r.sqr();
            if (e[i / 64] >> (i % 64) & 1) r = r * *this;
        }
        return r;
    }
    // lexicographic sign used by the compressed G2 encoding
    bool is_lex_largest() const { return c1.is_zero() ? c0.is_lex_largest() : c1.is_lex_largest(); }
    // square root, p = 3 mod 4; returns false if none
    bool sqrt(Fp2& out) const {
        if (is_zero()) { out = zero(); return true; }
        const u64* p = fp_params().mod;
        u64 e1[6], e2[6], three[6] = {3}, onev[6] = {1};
        sub_raw<6>(e1, p, three);                 // (p-3)/4
        for (int i = 0; i < 6; ++i) e1[i] = (e1[i] >> 2) | (i < 5 ? e1[i + 1] << 62 : 0);
        // re-do shift properly (the loop above reads already-shifted limbs only upward, which is fine)
        sub_raw<6>(e2, p, onev);                  // (p-1)/2
        for (int i = 0; i < 6; ++i) e2[i] = (e2[i] >> 1) | (i < 5 ? e2[i + 1] << 63 : 0);
        Fp2 a1 = pow(e1, 6);
        Fp2 alpha = a1.sqr() * *this;
        Fp2 x0 = a1 * *this;
        Fp2 x;
        if (alpha == -one()) {
            x = Fp2{Fp::zero(), Fp::one()} * x0;
        } else {
            Fp2 b = (one() + alpha).pow(e2, 6);
            x = b * x0;
        }
        if (x.sqr() != *this) return false;
        out = x;
        return true;
    }
};

struct Fp6 {
    Fp2 c0, c1, c2;
    static Fp6 zero() { return {Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
    static Fp6 one() { return {Fp2::one(), Fp2::zero(), Fp2::zero()}; }
    bool operator==(const Fp6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fp6 operator+(const Fp6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fp6 operator-(const Fp6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fp6 operator-() const { return {-c0, -c1, -c2}; }
    Fp6 operator*(const Fp6& o) const {    // schoolbook with v^3 = xi
        Fp2 t0 = c0 * o.c0 + (c1 * o.c2 + c2 * o.c1).mul_xi();
        Fp2 t1 = c0 * o.c1 + c1 * o.c0 + (c2 * o.c2).mul_xi();
        Fp2 t2 = c0 * o.c2 + c1 * o.c1 + c2 * o.c0;
        return {t0, t1, t2};
    }
    Fp6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
    Fp6 mul_fp2(const Fp2& s) const { return {c0 * s, c1 * s, c2 * s}; }
    Fp6 inv() const {
        Fp2 t0 = c0.sqr() - (c1 * c2).mul_xi();
        Fp2 t1 = c2.sqr().mul_xi() - c0 * c1;
        Fp2 t2 = c1.sqr() - c0 * c2;
        Fp2 d = (c0 * t0 + (c2 * t1 + c1 * t2).mul_xi()).inv();
        return {t0 * d, t1 * d, t2 * d};
    }
};

struct Fp12 {
    Fp6 c0, c1;
    static Fp12 one() { return {Fp6::one(), Fp6::zero()}; }
    bool operator==(const Fp12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fp12 operator*(const Fp12& o) const {
        Fp6 a = c0 * o.c0, b = c1 * o.c1;
        return {a + b.mul_v(), (c0 + c1) * (o.c0 + o.c1) - a - b};
    }
    Fp12 sqr() const { return *this * *this; }
    Fp12 conj() const { return {c0, -c1}; }
    Fp12 inv() const {
        Fp6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * d, -(c1 * d)};
    }
    // coefficient of w^k (k = 0..5) as Fp2: w^0,w^2,w^4 -> c0.{0,1,2}; w^1,w^3,w^5 -> c1.{0,1,2}
    Fp2& coef(int k) { Fp6& h = (k & 1) ? c1 : c0; return k / 2 == 0 ? h.c0 : (k / 2 == 1 ? h.c1 : h.c2); }
    const Fp2& coef(int k) const { return const_cast<Fp12*>(this)->coef(k); }
    Fp12 pow(const u64* e, int nl) const {
        Fp12 r = one();
        for (int i = nl * 64 - 1; i >= 0; --i) {
            r = r.sqr();
            if (e[i / 64] >> (i % 64) & 1) r = r * *this;
        }
        return r;
    }
};

// gamma[j][k] = xi^(k (p^j - 1)/6), j = 1,2 ; frob_j(sum a_k w^k) = sum conj^j(a_k) gamma[j][k] w^k
struct FrobTable {
    Fp2 g1[6], g2[6];
    FrobTable() {
        // (p-1)/6
        const u64* p = fp_params().mod;
        u64 e[6], onev[6] = {1};
        sub_raw<6>(e, p, onev);
        // divide by 6 (exact): long division by small constant
        u128 rem = 0;
        for (int i = 5; i >= 0; --i) {
            u128 cur = (rem << 64) | e[i];
            e[i] = (u64)(cur / 6);
            rem = cur % 6;
        }
        Fp2 xi{Fp::one(), Fp::one()};
        Fp2 base = xi.pow(e, 6);                       // xi^((p-1)/6) = w^(p-1)
        g1[0] = Fp2::one();
        for (int k = 1; k < 6; ++k) g1[k] = g1[k - 1] * base;
        // w^(p^2-1) = (w^(p-1))^(p+1) = base^p * base = conj(base)*base  (norm, in Fp)
        Fp2 base2 = base.conj() * base;
        g2[0] = Fp2::one();
        for (int k = 1; k < 6; ++k) g2[k] = g2[k - 1] * base2;
    }
};
inline const FrobTable& frob_table() {
    static const FrobTable t;
    return t;
}
inline Fp12 frob1(const Fp12& a) {
    Fp12 r;
    for (int k = 0; k < 6; ++k) r.coef(k) = a.coef(k).conj() * frob_table().g1[k];
    return r;
}
inline Fp12 frob2(const Fp12& a) {
    Fp12 r;
    for (int k = 0; k < 6; ++k) r.coef(k) = a.coef(k) * frob_table().g2[k];
    return r;
}

}  // namespace orc
