// Blob batch (include/kzgb200.h "Blob batch") -- device bodies.
//   leaf_k = SHA256("KZGB200/bleaf_v1" | 1 KiB of the blob), k < 128        (one thread per leaf)
//   z      = int_be(SHA256("KZGB200/blobz_v1" | C | leaf_0 .. leaf_127)) mod r   (one thread per blob)
//   y      = p(z) = (z^4096 - 1)/4096 * sum_i f_i w_i / (z - w_i),  w_i = w4096^brp12(i); f_i when z = w_i
//            (KZ_BLOB_THREADS lanes per blob, each with its own batch inversion over 4096/KZ_BLOB_THREADS terms)
// Domain points come from the context's table W[t] = omega_8192^-t:  w4096^e = W[(8192 - 2e) mod 8192].
#pragma once
#include "cells.cuh"

#define KZ_BLOB_LEN 4096
#define KZ_BLOB_LEAVES 128
#define KZ_BLOB_THREADS 128
#define KZ_BLOB_PER_LANE (KZ_BLOB_LEN / KZ_BLOB_THREADS)

KZ_HD u32 brp12(u32 v) { u32 r = 0; KZ_UNROLL for (int i = 0; i < 12; ++i) r |= ((v >> i) & 1u) << (11 - i); return r; }

// piece: 256 words holding 1024 big-endian bytes as stored in memory
KZ_HD void blob_leaf_words(u32 out[8], const u32* piece_le_words) {
    u32 h[8], w[16];
    sha256_init(h);
    const u32 total_words = 4 + 256;                         // 1040 bytes -> 16 full blocks + 4 words
    for (u32 blk = 0; blk < 17; ++blk) {
        for (int j = 0; j < 16; ++j) {
            u32 s = blk * 16 + j, v;
            if (s < 4) v = s == 0 ? TAGW('K', 'Z', 'G', 'B') : s == 1 ? TAGW('2', '0', '0', '/') : s == 2 ? TAGW('b', 'l', 'e', 'a') : TAGW('f', '_', 'v', '1');
            else if (s < total_words) v = bswap32(piece_le_words[s - 4]);
            else if (s == total_words) v = 0x80000000u;
            else v = 0;
            w[j] = v;
        }
        if (blk == 16) { w[14] = 0; w[15] = total_words * 32; }
        sha256_compress(h, w);
    }
    KZ_UNROLL for (int i = 0; i < 8; ++i) out[i] = h[i];
}
// C: 12 words of 48 big-endian bytes as stored; leaves: 128 x 8 digest words (values).  Returns z canonical (raw limbs).
KZ_HD Fr blob_z(const u32* C_le_words, const u32* leaves) {
    u32 h[8], w[16];
    sha256_init(h);
    const u32 total_words = 4 + 12 + 8 * KZ_BLOB_LEAVES;     // 1040 words = 65 blocks exactly, + 1 padding block
    for (u32 blk = 0; blk < 66; ++blk) {
        for (int j = 0; j < 16; ++j) {
            u32 s = blk * 16 + j, v;
            if (s < 4) v = s == 0 ? TAGW('K', 'Z', 'G', 'B') : s == 1 ? TAGW('2', '0', '0', '/') : s == 2 ? TAGW('b', 'l', 'o', 'b') : TAGW('z', '_', 'v', '1');
            else if (s < 16) v = bswap32(C_le_words[s - 4]);
            else if (s < total_words) v = leaves[s - 16];
            else if (s == total_words) v = 0x80000000u;
            else v = 0;
            w[j] = v;
        }
        if (blk == 65) { w[14] = 0; w[15] = total_words * 32; }
        sha256_compress(h, w);
    }
    Fr raw;
    KZ_UNROLL for (int i = 0; i < 8; ++i) raw.v[i] = h[7 - i];
    return fr_reduce_raw(raw);                               // 2^256 < 3r: two conditional subtractions
}

struct BlobLane { Fr sum; Fr hit; u32 has_hit; u32 bad; };
// lane `lane` of KZ_BLOB_THREADS: its share sum_i f_i w_i / (z - w_i) over i = lane + j * KZ_BLOB_THREADS.
// z in Montgomery form.  blob: 4096 x 32 big-endian bytes.
KZ_COLD BlobLane blob_eval_lane(const Fr* W, const u8* blob, const Fr& z, u32 lane) {
    BlobLane L;
    L.sum = fr_zero(); L.hit = fr_zero(); L.has_hit = 0; L.bad = 0;
    Fr pre[KZ_BLOB_PER_LANE];
    Fr acc = fr_const(FR_ONE);
    for (int j = 0; j < KZ_BLOB_PER_LANE; ++j) {             // prefix products of the denominators
        u32 i = lane + (u32)j * KZ_BLOB_THREADS;
        Fr d = fr_sub(z, W[(KZ_N_EXT - 2u * brp12(i)) & (KZ_N_EXT - 1u)]);
        if (fr_is_zero(d)) d = fr_const(FR_ONE);
        pre[j] = acc;
        acc = fr_mul(acc, d);
    }
    Fr inv = fr_inv(acc);
    for (int j = KZ_BLOB_PER_LANE - 1; j >= 0; --j) {
        u32 i = lane + (u32)j * KZ_BLOB_THREADS;
        Fr wi = W[(KZ_N_EXT - 2u * brp12(i)) & (KZ_N_EXT - 1u)];
        Fr d = fr_sub(z, wi);
        Fr f;
        fr_raw_from_be(f, blob + 32 * i);
        if (!fr_raw_is_canonical(f)) { L.bad += 1; f = fr_zero(); }
        f = fr_to_mont(f);
        if (fr_is_zero(d)) { L.has_hit = 1; L.hit = f; d = fr_const(FR_ONE); }
        Fr di = fr_mul(inv, pre[j]);
        inv = fr_mul(inv, d);
        L.sum = fr_add(L.sum, fr_mul(fr_mul(f, wi), di));
    }
    return L;
}
// y from the total of the lane sums (Montgomery in, Montgomery out)
KZ_COLD Fr blob_eval_finish(const Fr& z, const Fr& total) {
    Fr zn = z;
    for (int i = 0; i < 12; ++i) zn = fr_mul(zn, zn);
    zn = fr_sub(zn, fr_const(FR_ONE));
    return fr_mul(fr_mul(zn, fr_const(FR_INV4096)), total);
}
