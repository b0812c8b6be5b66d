// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Scalar SHA-256 (FIPS 180-4), used for the batched
// Fiat-Shamir construction of SURVEY.md App. B.4 and the synthetic-input PRNG.
// Round constants are computed from the cube roots of the first 64 primes at start-up.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;

    static const uint32_t* K() {
        static uint32_t k[64];
        static bool init = [] {
            int n = 0;
            for (int c = 2; n < 64; ++c) {
                bool prime = true;
                for (int d = 2; d * d <= c; ++d) if (c % d == 0) prime = false;
                if (!prime) continue;
                long double r = cbrtl((long double)c);
                r -= floorl(r);
                k[n++] = (uint32_t)floorl(r * 4294967296.0L);
            }
            return true;
        }();
        (void)init;
        return k;
    }
    Sha256() {
        static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                       0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        memcpy(h, iv, sizeof h);
    }
    static uint32_t rotr(uint32_t x, int n) { return x >> n | x << (32 - n); }
    void block(const uint8_t* p) {
        const uint32_t* k = K();
        uint32_t w[64];
        for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] << 24 | p[4 * i + 1] << 16 | p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + k[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data;
        size_t fill = len % 64;
        len += n;
        if (fill) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            p += take; n -= take;
            if (fill + take < 64) return;
            block(buf);
        }
        for (; n >= 64; p += 64, n -= 64) block(p);
        if (n) memcpy(buf, p, n);
    }
    void update_u64be(uint64_t v) {
        uint8_t b[8];
        for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(v >> (56 - 8 * i));
        update(b, 8);
    }
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad[72] = {0x80};
        size_t padlen = (len % 64 < 56 ? 56 : 120) - len % 64;
        update(pad, padlen);
        uint8_t lb[8];
        for (int i = 0; i < 8; ++i) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; ++i) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
    }
};

}  // namespace orc
