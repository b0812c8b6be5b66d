// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Blob-level caller of the batch check:
// blob -> evaluation challenge z by hashing, y = p(z) by barycentric evaluation, then the plain KZG batch.
// EIP-4844-shaped, conventions are this repo's (DESIGN.md "Blob batch"); the upstream reference is LICENSE-only
// (/root/reference/LICENSE:1-201), semantics pinned by tests against direct polynomial evaluation (oracle/pymodel).
#pragma once
#include "cells.hpp"

namespace orc {

constexpr size_t BLOB_LEN = 4096, BLOB_BYTES = 4096 * 32, BLOB_LEAVES = 128, BLOB_LEAF_BYTES = 1024;
constexpr u64 STREAM_BLOBEV = 12;
static const char TAG_BLEAF[] = "KZGB200/bleaf_v1";
static const char TAG_BLOBZ[] = "KZGB200/blobz_v1";

// evaluation domain in blob order: blob[i] = p(w4096^brp12(i)), w4096 = omega_ext^2
inline const std::vector<Fr>& blob_domain() {
    static const std::vector<Fr> d = [] {
        std::vector<Fr> pw(BLOB_LEN), out(BLOB_LEN);
        Fr w = omega_ext() * omega_ext(), t = Fr::one();
        for (size_t e = 0; e < BLOB_LEN; ++e) { pw[e] = t; t = t * w; }
        for (size_t i = 0; i < BLOB_LEN; ++i) out[i] = pw[brp((unsigned)i, 12)];
        return out;
    }();
    return d;
}
// z = int_be(SHA256("KZGB200/blobz_v1" | C | leaf_0 .. leaf_127)) mod r,  leaf_k = SHA256("KZGB200/bleaf_v1" | 1 KiB of the blob)
inline Fr blob_challenge(const u8* blob, const u8* C) {
    Sha256 top;
    top.update(TAG_BLOBZ, 16);
    top.update(C, 48);
    for (size_t k = 0; k < BLOB_LEAVES; ++k) {
        Sha256 s;
        s.update(TAG_BLEAF, 16);
        s.update(blob + BLOB_LEAF_BYTES * k, BLOB_LEAF_BYTES);
        u8 d[32];
        s.final(d);
        top.update(d, 32);
    }
    u8 d[32], wide[64] = {0};
    top.final(d);
    memcpy(wide + 32, d, 32);
    return fr_from_512(wide);
}
// y = p(z) from the evaluations; returns the number of blob elements >= r (y is then meaningless)
inline unsigned blob_eval(Fr& y, const u8* blob, const Fr& z) {
    const std::vector<Fr>& dom = blob_domain();
    std::vector<Fr> f(BLOB_LEN), den(BLOB_LEN);
    unsigned bad = 0;
    long hit = -1;
    for (size_t i = 0; i < BLOB_LEN; ++i) {
        if (!fr_from_be(f[i], blob + 32 * i)) { ++bad; f[i] = Fr::zero(); }
        den[i] = z - dom[i];
        if (den[i].is_zero()) { hit = (long)i; den[i] = Fr::one(); }
    }
    if (bad) { y = Fr::zero(); return bad; }
    if (hit >= 0) { y = f[(size_t)hit]; return 0; }
    // batch inversion
    std::vector<Fr> pre(BLOB_LEN);
    Fr acc = Fr::one();
    for (size_t i = 0; i < BLOB_LEN; ++i) { pre[i] = acc; acc = acc * den[i]; }
    Fr inv = acc.inv(), sum = Fr::zero();
    for (size_t i = BLOB_LEN; i-- > 0;) {
        Fr di = inv * pre[i];
        inv = inv * den[i];
        sum = sum + f[i] * dom[i] * di;
    }
    Fr zn = fr_pow_u64(z, BLOB_LEN) - Fr::one();
    y = zn * Fr::from_u64(BLOB_LEN).inv() * sum;
    return 0;
}
// z and y (32 B big-endian each) of every blob; returns the number of malformed blob elements
inline unsigned blob_challenges_evals(u8* z_out, u8* y_out, const u8* blobs, const u8* comms, size_t m, int threads) {
    std::vector<unsigned> bad(m, 0);
    parallel_for(m, threads, [&](size_t b0, size_t b1) {
        for (size_t j = b0; j < b1; ++j) {
            Fr z = blob_challenge(blobs + BLOB_BYTES * j, comms + 48 * j), y;
            bad[j] = blob_eval(y, blobs + BLOB_BYTES * j, z);
            z.to_bytes_be(z_out + 32 * j);
            y.to_bytes_be(y_out + 32 * j);
        }
    });
    unsigned tot = 0;
    for (unsigned v : bad) tot += v;
    return tot;
}
// synthetic blobs with the known test tau: random evaluations, C = [p(tau)]G1, proof = [(p(tau) - y)/(tau - z)]G1
inline void synth_blobs(u64 seed, size_t m, u8* blobs, u8* comms, u8* proofs, int threads) {
    const FixedBase& fb = g1_fixed_base();
    Fr tau = test_tau();
    parallel_for(m, threads, [&](size_t b0, size_t b1) {
        for (size_t j = b0; j < b1; ++j) {
            u8* blob = blobs + BLOB_BYTES * j;
            for (size_t i = 0; i < BLOB_LEN; ++i) prng_fr(seed, STREAM_BLOBEV, j * BLOB_LEN + i).to_bytes_be(blob + 32 * i);
            Fr pt, y;
            blob_eval(pt, blob, tau);
            g1_compress(comms + 48 * j, g1_affine(fb.mul(pt)));
            Fr z = blob_challenge(blob, comms + 48 * j);
            blob_eval(y, blob, z);
            g1_compress(proofs + 48 * j, g1_affine(fb.mul((pt - y) * (tau - z).inv())));
        }
    });
}

}  // namespace orc
