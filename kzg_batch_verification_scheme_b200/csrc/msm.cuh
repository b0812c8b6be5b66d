// Pippenger multi-scalar multiplication over G1 -- per-thread bodies (BASELINE.json:5 item (d)):
// signed-digit windows, (bucket, point) pairs sorted by bucket, bucket accumulation in XYZZ, parallel
// bucket reduction, window combine.  The __global__ wrappers and the sort live in msm.cu.
#pragma once
#include "g1.cuh"

#define KZ_MSM_MAX_WINDOWS 96
#define KZ_MSM_SEG 32            // buckets per reduction segment

struct MsmPlan {
    int nbits;                   // scalar width: 128 or 255
    int c;                       // window width
    int W;                       // number of windows = ceil(nbits / c)
    u32 nb[KZ_MSM_MAX_WINDOWS];          // buckets per window (top window is unsigned)
    u32 bucket_off[KZ_MSM_MAX_WINDOWS + 1];   // prefix sums of nb
    u32 seg_off[KZ_MSM_MAX_WINDOWS + 1];      // prefix sums of ceil(nb / SEG)
    u32 total_buckets, total_segs;
};

// Signed windows 0..W-2 (digit in [-2^(c-1), 2^(c-1)]), unsigned top window that absorbs the last carry,
// so no extra window is ever needed: top window holds tb = nbits - c(W-1) bits, values 0..2^tb.
inline MsmPlan msm_make_plan(size_t n, int nbits) {
    MsmPlan p;
    int lg = 0;
    while ((size_t(1) << (lg + 1)) <= n) ++lg;
    int c = lg - 3;
    if (c < 3) c = 3;
    if (c > 16) c = 16;
    p.nbits = nbits;
    p.c = c;
    p.W = (nbits + c - 1) / c;
    p.bucket_off[0] = 0;
    p.seg_off[0] = 0;
    for (int w = 0; w < p.W; ++w) {
        int tb = nbits - c * (p.W - 1);
        p.nb[w] = (w < p.W - 1) ? (1u << (c - 1)) : (1u << tb);
        p.bucket_off[w + 1] = p.bucket_off[w] + p.nb[w];
        p.seg_off[w + 1] = p.seg_off[w] + (p.nb[w] + KZ_MSM_SEG - 1) / KZ_MSM_SEG;
    }
    p.total_buckets = p.bucket_off[p.W];
    p.total_segs = p.seg_off[p.W];
    return p;
}

// digits of one scalar (8 little-endian limbs, only nbits significant).  Writes, for every window w,
// key = global bucket id (or total_buckets for a zero digit) and val = point index | sign << 31 at
// position w*n + i.
KZ_HD void msm_digits_body(u32* keys, u32* vals, const u32* sc, size_t i, size_t n, const MsmPlan& P) {
    u32 carry = 0;
    const int c = P.c;
    for (int w = 0; w < P.W; ++w) {
        int lo = w * c;
        int width = (w < P.W - 1) ? c : (P.nbits - lo);
        u32 limb = lo >> 5, sh = lo & 31;
        u64 two = sc[limb];
        if (limb + 1 < 8) two |= (u64)sc[limb + 1] << 32;
        u32 v = (u32)(two >> sh) & ((width == 32) ? 0xFFFFFFFFu : ((1u << width) - 1u));
        v += carry;
        u32 neg = 0;
        if (w < P.W - 1 && v > (1u << (c - 1))) { v = (1u << c) - v; neg = 1; carry = 1; } else carry = 0;
        keys[(size_t)w * n + i] = v ? (P.bucket_off[w] + v - 1) : P.total_buckets;
        vals[(size_t)w * n + i] = (u32)i | (neg << 31);
    }
}

KZ_HD G1Aff load_point(const Fp* pts, size_t idx) { return {pts[2 * idx], pts[2 * idx + 1]}; }

// one bucket: sum of +-P over the sorted range [lo, hi)
KZ_HD G1Xyzz msm_bucket_body(const Fp* pts, const u32* vals, u32 lo, u32 hi) {
    G1Xyzz acc = xyzz_inf();
    for (u32 j = lo; j < hi; ++j) {
        u32 v = vals[j];
        G1Aff p = load_point(pts, v & 0x7FFFFFFFu);
        if (aff_is_inf(p)) continue;
        if (v >> 31) p.y = fp_neg(p.y);
        acc = xyzz_madd(acc, p);
    }
    return acc;
}

// [k]P by double-and-add for a small k (< 2^17)
KZ_COLD G1Xyzz xyzz_mul_small(const G1Xyzz& p, u32 k) {
    G1Xyzz r = xyzz_inf();
    for (int i = 16; i >= 0; --i) {
        r = xyzz_dbl(r);
        if ((k >> i) & 1) r = xyzz_add(r, p);
    }
    return r;
}
// segment `seg` of window w: local buckets k = base+1 .. min(base+SEG, nb), base = seg*SEG.
// returns sum_k k * B_k  (running-sum trick inside the segment plus [base] * run).
KZ_HD G1Xyzz msm_segment_body(const G1Xyzz* buckets, u32 nb, u32 seg) {
    u32 base = seg * KZ_MSM_SEG;
    u32 top = base + KZ_MSM_SEG < nb ? base + KZ_MSM_SEG : nb;
    G1Xyzz run = xyzz_inf(), acc = xyzz_inf();
    for (u32 k = top; k > base; --k) {          // bucket with value k is stored at index k-1
        run = xyzz_add(run, buckets[k - 1]);
        acc = xyzz_add(acc, run);
    }
    if (base) acc = xyzz_add(acc, xyzz_mul_small(run, base));
    return acc;
}
// Horner over window sums: result = sum_w 2^(c w) * win[w]
KZ_COLD G1Xyzz msm_combine_body(const G1Xyzz* win, int W, int c) {
    G1Xyzz acc = xyzz_inf();
    for (int w = W - 1; w >= 0; --w) {
        for (int k = 0; k < c; ++k) acc = xyzz_dbl(acc);
        acc = xyzz_add(acc, win[w]);
    }
    return acc;
}
