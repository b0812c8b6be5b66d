// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
// Prime-field arithmetic in 64-bit limbs, Montgomery form, scalar code with unsigned __int128.
// Deliberately independent of the CUDA path (which uses 32-bit limbs and generated constants):
// every Montgomery constant here is computed at start-up from the modulus alone.
//
// Reference citation: upstream /root/reference holds only LICENSE:1-201; semantics follow
// BASELINE.json:5 ("scalar, multithreaded C++ CPU oracle") and SURVEY.md 2.2 row S9.
// Parity pin: constants and results are checked against oracle/pymodel (pure Python) and the
// known answers in SURVEY.md Appendix A; "reference parity unpinned" (nothing upstream to pin to).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace orc {
using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using u128 = unsigned __int128;

template <int N>
struct Params {
    u64 mod[N];
    u64 ninv;      // -mod^-1 mod 2^64
    u64 one[N];    // R mod p
    u64 r2[N];     // R^2 mod p
    u64 half[N];   // (p-1)/2, canonical
};

template <int N>
inline bool ge_raw(const u64* a, const u64* b) {
    for (int i = N - 1; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
template <int N>
inline u64 add_raw(u64* r, const u64* a, const u64* b) {
    u64 c = 0;
    for (int i = 0; i < N; ++i) {
        u128 t = (u128)a[i] + b[i] + c;
        r[i] = (u64)t;
        c = (u64)(t >> 64);
    }
    return c;
}
template <int N>
inline u64 sub_raw(u64* r, const u64* a, const u64* b) {
    u64 bw = 0;
    for (int i = 0; i < N; ++i) {
        u128 t = (u128)a[i] - b[i] - bw;
        r[i] = (u64)t;
        bw = (u64)(t >> 64) & 1;
    }
    return bw;
}

template <int N>
Params<N> make_params(const char* hex) {
    Params<N> p{};
    int len = (int)strlen(hex);
    for (int i = 0; i < len; ++i) {
        char ch = hex[len - 1 - i];
        u64 v = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
        p.mod[i / 16] |= v << (4 * (i % 16));
    }
    u64 inv = 1;                                   // Newton: inv = mod^-1 mod 2^64
    for (int i = 0; i < 6; ++i) inv *= 2 - p.mod[0] * inv;
    p.ninv = 0 - inv;
    u64 t[N] = {1};                                // t = 2^k mod p by doubling
    auto dbl = [&]() {
        u64 c = add_raw<N>(t, t, t);
        if (c || ge_raw<N>(t, p.mod)) sub_raw<N>(t, t, p.mod);
    };
    for (int i = 0; i < 64 * N; ++i) dbl();
    memcpy(p.one, t, sizeof t);
    for (int i = 0; i < 64 * N; ++i) dbl();
    memcpy(p.r2, t, sizeof t);
    u64 pm1[N];
    u64 onev[N] = {1};
    sub_raw<N>(pm1, p.mod, onev);
    for (int i = 0; i < N; ++i) p.half[i] = (pm1[i] >> 1) | (i + 1 < N ? pm1[i + 1] << 63 : 0);
    return p;
}

template <int N, const Params<N>& (*PP)()>
struct Fe {
    u64 l[N];
    static constexpr int LIMBS = N;
    static constexpr int BYTES = 8 * N;

    static Fe zero() { Fe r; memset(r.l, 0, sizeof r.l); return r; }
    static Fe one() { Fe r; memcpy(r.l, PP().one, sizeof r.l); return r; }
    static Fe from_u64(u64 v) { u64 raw[N] = {v}; return from_raw(raw); }
    // raw canonical limbs (little-endian u64, value < p) -> Montgomery
    static Fe from_raw(const u64* raw) {
        Fe a, r2;
        memcpy(a.l, raw, sizeof a.l);
        memcpy(r2.l, PP().r2, sizeof r2.l);
        return a * r2;
    }
    void to_raw(u64* raw) const {
        Fe o = zero();
        o.l[0] = 1;
        Fe c = *this * o;
        memcpy(raw, c.l, sizeof c.l);
    }
    // big-endian bytes; returns false if value >= p (result then undefined)
    static bool from_bytes_be(Fe& out, const u8* b) {
        u64 raw[N];
        for (int i = 0; i < N; ++i) {
            u64 v = 0;
            for (int k = 0; k < 8; ++k) v = v << 8 | b[8 * (N - 1 - i) + k];
            raw[i] = v;
        }
        if (ge_raw<N>(raw, PP().mod)) return false;
        out = from_raw(raw);
        return true;
    }
    void to_bytes_be(u8* b) const {
        u64 raw[N];
        to_raw(raw);
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < 8; ++k) b[8 * (N - 1 - i) + k] = (u8)(raw[i] >> (56 - 8 * k));
    }
    bool is_zero() const {
        u64 a = 0;
        for (int i = 0; i < N; ++i) a |= l[i];
        return a == 0;
    }
    bool operator==(const Fe& o) const { return memcmp(l, o.l, sizeof l) == 0; }
    bool operator!=(const Fe& o) const { return !(*this == o); }
    // canonical value > (p-1)/2  ("lexicographically largest")
    bool is_lex_largest() const {
        u64 raw[N];
        to_raw(raw);
        return !ge_raw<N>(PP().half, raw);
    }
    Fe operator+(const Fe& o) const {
        Fe r;
        u64 c = add_raw<N>(r.l, l, o.l);
        if (c || ge_raw<N>(r.l, PP().mod)) sub_raw<N>(r.l, r.l, PP().mod);
        return r;
    }
    Fe operator-(const Fe& o) const {
        Fe r;
        if (sub_raw<N>(r.l, l, o.l)) add_raw<N>(r.l, r.l, PP().mod);
        return r;
    }
    Fe operator-() const { return zero() - *this; }
    Fe dbl() const { return *this + *this; }
    // CIOS Montgomery multiplication
    Fe operator*(const Fe& o) const {
        const u64* m = PP().mod;
        const u64 ninv = PP().ninv;
        u64 t[N + 2] = {0};
        for (int i = 0; i < N; ++i) {
            u64 c = 0;
            for (int j = 0; j < N; ++j) {
                u128 v = (u128)l[j] * o.l[i] + t[j] + c;
                t[j] = (u64)v;
                c = (u64)(v >> 64);
            }
            u128 v = (u128)t[N] + c;
            t[N] = (u64)v;
            t[N + 1] = (u64)(v >> 64);
            u64 q = t[0] * ninv;
            v = (u128)q * m[0] + t[0];
            c = (u64)(v >> 64);
            for (int j = 1; j < N; ++j) {
                v = (u128)q * m[j] + t[j] + c;
                t[j - 1] = (u64)v;
                c = (u64)(v >> 64);
            }
            v = (u128)t[N] + c;
            t[N - 1] = (u64)v;
            t[N] = t[N + 1] + (u64)(v >> 64);
        }
        Fe r;
        memcpy(r.l, t, sizeof r.l);
        if (t[N] || ge_raw<N>(r.l, m)) sub_raw<N>(r.l, r.l, m);
        return r;
    }
    Fe sqr() const { return *this * *this; }
    // exponent as little-endian u64 limbs
    Fe pow(const u64* e, int nl) const {
        Fe r = one();
        bool started = false;
        for (int i = nl * 64 - 1; i >= 0; --i) {
            if (started) r = r.sqr();
            if (e[i / 64] >> (i % 64) & 1) {
                r = started ? r * *this : *this;
                started = true;
            }
        }
        return r;
    }
    Fe inv() const {   // Fermat; inv(0) = 0
        u64 e[N];
        u64 two[N] = {2};
        sub_raw<N>(e, PP().mod, two);
        return pow(e, N);
    }
};

// ---- concrete fields
inline const Params<6>& fp_params() {
    static const Params<6> p = make_params<6>(
        "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab");
    return p;
}
inline const Params<4>& fr_params() {
    static const Params<4> p = make_params<4>("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001");
    return p;
}
using Fp = Fe<6, fp_params>;
using Fr = Fe<4, fr_params>;

}  // namespace orc
