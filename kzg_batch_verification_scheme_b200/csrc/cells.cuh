// Cell batch (BASELINE.json config[4]) -- device bodies.  One block of 64 threads per
// opening: range-check the 64 evaluations, inverse NTT of size 64 in shared memory, scale coefficient i by
// h_c^-i / 64 (coset shift) and by the challenge r_k.  Twiddles come from one table W[t] = omega^-t, t < 8192
// (omega_64 = omega^128, h_c = omega^brp7(c)), built once per context.
#pragma once
#include "pairing.cuh"      // COOP_FOR / COOP_SYNC
#include "sha256.cuh"

#define KZ_N_EXT 8192u
#define KZ_CELL_LEN 64
#define KZ_CELL_WORDS 512      // 2048 bytes

KZ_HD u32 brp7(u32 c) { c &= 127u; u32 r = 0; KZ_UNROLL for (int i = 0; i < 7; ++i) r |= ((c >> i) & 1u) << (6 - i); return r; }
KZ_HD u32 brp6(u32 c) { u32 r = 0; KZ_UNROLL for (int i = 0; i < 6; ++i) r |= ((c >> i) & 1u) << (5 - i); return r; }

// leaf_k = SHA256("KZGB200/cell_v1_" | u64be(commitment index) | u64be(cell index) | cell (2048 B) | proof (48 B))
// cell / proof given as pointers to big-endian bytes (word-aligned); 532 message words -> 34 blocks
KZ_HD void fs_cell_leaf_words(u32 out[8], u32 ci, u32 xi, const u32* cell_le_words, const u32* proof_le_words) {
    u32 h[8], w[16];
    sha256_init(h);
    const u32 total_words = 4 + 4 + KZ_CELL_WORDS + 12;      // 532
    const u32 nblocks = total_words / 16 + 1;                 // 34 (532 = 33*16 + 4)
    for (u32 blk = 0; blk < nblocks; ++blk) {
        for (int j = 0; j < 16; ++j) {
            u32 s = blk * 16 + j, v;
            if (s < 4) { v = s == 0 ? TAGW('K', 'Z', 'G', 'B') : s == 1 ? TAGW('2', '0', '0', '/') : s == 2 ? TAGW('c', 'e', 'l', 'l') : TAGW('_', 'v', '1', '_'); }
            else if (s < 8) v = s == 5 ? ci : (s == 7 ? xi : 0u);
            else if (s < 8 + KZ_CELL_WORDS) v = bswap32(cell_le_words[s - 8]);
            else if (s < total_words) v = bswap32(proof_le_words[s - 8 - KZ_CELL_WORDS]);
            else if (s == total_words) v = 0x80000000u;
            else v = 0;
            w[j] = v;
        }
        if (blk == nblocks - 1) { w[14] = 0; w[15] = total_words * 32; }
        sha256_compress(h, w);
    }
    KZ_UNROLL for (int i = 0; i < 8; ++i) out[i] = h[i];
}

struct CellScratch {
    Fr a[KZ_CELL_LEN];
    Fr rk;          // challenge, Montgomery form
    u32 bad;
};

// One opening (64 cooperating threads).  cell: 2048 big-endian bytes.  Outputs:
//   coef_out[64]: r_k * a_i (Montgomery), a = interpolation coefficients of the cell
//   r_out[4]:     r_k (canonical limbs);   rh_out[8]: r_k * h_c^64 (canonical limbs)
// returns the number of malformed items seen by the calling thread set (summed in S.bad).
KZ_COLD void coop_cell_body(CellScratch& S, const Fr* W, const u32* root_words, u64 k, u32 ci, u32 xi, u32 nc, const u8* cell,
                            Fr* coef_out, u32* r_out, u32* rh_out) {
    COOP_FOR(t, 1) {
        S.bad = (xi >= 128u ? 1u : 0u) + (ci >= nc ? 1u : 0u);
        Fr r = fr_zero();
        u32 root[8];
        for (int i = 0; i < 8; ++i) root[i] = root_words[i];
        fs_r_limbs(r.v, root, k);
        for (int i = 0; i < 4; ++i) r_out[i] = r.v[i];
        S.rk = fr_to_mont(r);
        // r_k * h^64 with h^64 = omega^(64 brp7(c)) = W[8192 - 64 brp7(c)]
        u32 e = (KZ_N_EXT - 64u * brp7(xi)) & (KZ_N_EXT - 1u);
        Fr rh = fr_mul(r, W[e]);                 // raw * Montgomery -> canonical product
        for (int i = 0; i < 8; ++i) rh_out[i] = rh.v[i];
    }
    COOP_SYNC();
    // load in bit-reversed order (decimation in time), range check
    COOP_FOR(t, KZ_CELL_LEN) {
        Fr y;
        fr_raw_from_be(y, cell + 32 * t);
        bool okv = fr_raw_is_canonical(y);
        if (!okv) {
#if defined(KZGB_EMU)
            S.bad += 1;
#else
            atomicAdd(&S.bad, 1u);
#endif
            y = fr_zero();
        }
        S.a[brp6((u32)t)] = fr_to_mont(y);
    }
    COOP_SYNC();
    // inverse NTT: 6 radix-2 stages, 32 butterflies each; twiddle omega_64^-(k * 64/len) = W[128 * k * 64 / len]
    for (int len = 2; len <= KZ_CELL_LEN; len <<= 1) {
        COOP_FOR(t, KZ_CELL_LEN / 2) {
            int half = len >> 1, grp = t / half, kk = t - grp * half, i0 = grp * len + kk;
            Fr u = S.a[i0], v = fr_mul(S.a[i0 + half], W[(128u * (u32)kk * (u32)(KZ_CELL_LEN / len)) & (KZ_N_EXT - 1u)]);
            S.a[i0] = fr_add(u, v);
            S.a[i0 + half] = fr_sub(u, v);
        }
        COOP_SYNC();
    }
    // a_i = x_i / 64 * h^-i, then times r_k
    COOP_FOR(t, KZ_CELL_LEN) {
        Fr s = fr_mul(fr_const(FR_INV64), W[(brp7(xi) * (u32)t) & (KZ_N_EXT - 1u)]);
        coef_out[t] = fr_mul(fr_mul(S.a[t], s), S.rk);
    }
    COOP_SYNC();
}
