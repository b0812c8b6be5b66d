// SHA-256 compression on the ALU pipe and the batched Fiat-Shamir construction of BASELINE.json:5
// item (c) (layout: DESIGN.md "SPEC").  Messages are handled as big-endian words.
#pragma once
#include "field.cuh"

KZ_CONSTANT u32 SHA_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

KZ_HD u32 sha_rotr(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
KZ_HD u32 bswap32(u32 x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); }

KZ_HD void sha256_init(u32 h[8]) {
    h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
    h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
}
// one compression; w[16] is consumed (used as the rolling schedule)
KZ_HD void sha256_compress(u32 h[8], u32 w[16]) {
    u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    KZ_UNROLL for (int i = 0; i < 64; ++i) {
        if (i >= 16) {
            u32 w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            u32 s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            u32 s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
        }
        u32 t1 = hh + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i & 15];
        u32 t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// 16-byte domain tags as big-endian words
#define TAGW(a, b, c, d) (((u32)(a) << 24) | ((u32)(b) << 16) | ((u32)(c) << 8) | (u32)(d))
KZ_HD void tag_leaf(u32* w) {   // "KZGB200/leaf_v1_"
    w[0] = TAGW('K', 'Z', 'G', 'B'); w[1] = TAGW('2', '0', '0', '/'); w[2] = TAGW('l', 'e', 'a', 'f'); w[3] = TAGW('_', 'v', '1', '_');
}
KZ_HD void tag_chunk(u32* w) {  // "KZGB200/chunk_v1"
    w[0] = TAGW('K', 'Z', 'G', 'B'); w[1] = TAGW('2', '0', '0', '/'); w[2] = TAGW('c', 'h', 'u', 'n'); w[3] = TAGW('k', '_', 'v', '1');
}
KZ_HD void tag_r(u32* w) {      // "KZGB200/r_v1____"
    w[0] = TAGW('K', 'Z', 'G', 'B'); w[1] = TAGW('2', '0', '0', '/'); w[2] = TAGW('r', '_', 'v', '1'); w[3] = TAGW('_', '_', '_', '_');
}

// leaf_i = SHA256(tag | C_i | z_i | y_i | pi_i): 176 bytes -> 3 blocks.  Inputs as big-endian words.
KZ_HD void fs_leaf_words(u32 out[8], const u32 c[12], const u32 z[8], const u32 y[8], const u32 pi[12]) {
    u32 h[8], w[16];
    sha256_init(h);
    tag_leaf(w);
    KZ_UNROLL for (int i = 0; i < 12; ++i) w[4 + i] = c[i];
    sha256_compress(h, w);
    KZ_UNROLL for (int i = 0; i < 8; ++i) { w[i] = z[i]; w[8 + i] = y[i]; }
    sha256_compress(h, w);
    KZ_UNROLL for (int i = 0; i < 12; ++i) w[i] = pi[i];
    w[12] = 0x80000000u; w[13] = 0; w[14] = 0; w[15] = 176 * 8;
    sha256_compress(h, w);
    KZ_UNROLL for (int i = 0; i < 8; ++i) out[i] = h[i];
}
// chunk_j = SHA256(tag | leaf words of `nleaves` leaves); leaves = contiguous [nleaves][8] words
KZ_HD void fs_chunk_words(u32 out[8], const u32* leaves, u32 nleaves) {
    u32 h[8], w[16];
    sha256_init(h);
    u32 total_words = 4 + 8 * nleaves;              // always = 4 mod 8
    u32 nblocks = total_words / 16 + 1;
    for (u32 blk = 0; blk < nblocks; ++blk) {
        for (int j = 0; j < 16; ++j) {
            u32 s = blk * 16 + j;                    // word index in the message
            u32 v;
            if (s < 4) { u32 t[4]; tag_chunk(t); v = t[s]; }
            else if (s < total_words) v = leaves[s - 4];
            else if (s == total_words) v = 0x80000000u;
            else v = 0;
            w[j] = v;
        }
        if (blk == nblocks - 1) { w[14] = 0; w[15] = total_words * 32; }   // bit length (< 2^32)
        sha256_compress(h, w);
    }
    KZ_UNROLL for (int i = 0; i < 8; ++i) out[i] = h[i];
}
// r_i = first 16 bytes of SHA256(tag | root | u64be(i)) as 4 little-endian limbs (128-bit integer)
KZ_HD void fs_r_limbs(u32 r[4], const u32 root[8], u64 idx) {
    u32 h[8], w[16];
    sha256_init(h);
    tag_r(w);
    KZ_UNROLL for (int i = 0; i < 8; ++i) w[4 + i] = root[i];
    w[12] = (u32)(idx >> 32); w[13] = (u32)idx; w[14] = 0x80000000u; w[15] = 0;
    sha256_compress(h, w);
    KZ_UNROLL for (int i = 0; i < 14; ++i) w[i] = 0;
    w[14] = 0; w[15] = 56 * 8;
    sha256_compress(h, w);
    r[0] = h[3]; r[1] = h[2]; r[2] = h[1]; r[3] = h[0];
}
// synthetic-input PRNG block: SHA256("kzgb200/prng" | u64be(seed) | u64be(stream) | u64be(k)), 36 bytes
KZ_HD void prng_block_words(u32 out[8], u64 seed, u64 stream, u64 k) {
    u32 h[8], w[16];
    sha256_init(h);
    w[0] = TAGW('k', 'z', 'g', 'b'); w[1] = TAGW('2', '0', '0', '/'); w[2] = TAGW('p', 'r', 'n', 'g');
    w[3] = (u32)(seed >> 32); w[4] = (u32)seed; w[5] = (u32)(stream >> 32); w[6] = (u32)stream;
    w[7] = (u32)(k >> 32); w[8] = (u32)k; w[9] = 0x80000000u;
    KZ_UNROLL for (int i = 10; i < 15; ++i) w[i] = 0;
    w[15] = 36 * 8;
    sha256_compress(h, w);
    KZ_UNROLL for (int i = 0; i < 8; ++i) out[i] = h[i];
}
// Fr sample = (block(2i) | block(2i+1)) as a 512-bit big-endian integer mod r  -> Montgomery form
KZ_HD Fr prng_fr(u64 seed, u64 stream, u64 idx) {
    u32 hi[8], lo[8];
    prng_block_words(hi, seed, stream, 2 * idx);
    prng_block_words(lo, seed, stream, 2 * idx + 1);
    Fr a, b;
    KZ_UNROLL for (int i = 0; i < 8; ++i) { a.v[i] = hi[7 - i]; b.v[i] = lo[7 - i]; }
    // a*2^256 + b with both halves first reduced below r
    return fr_add(fr_mul(fr_to_mont(fr_reduce_raw(a)), fr_const(FR_2_256)), fr_to_mont(fr_reduce_raw(b)));
}
