// Pippenger multi-scalar multiplication over G1 -- per-thread bodies (BASELINE.json:5 item (d)):
// signed-digit windows, (bucket, point) pairs sorted by bucket, bucket accumulation in XYZZ, parallel
// bucket reduction, window combine.  The __global__ wrappers and the sort live in msm.cu.
#pragma once
#include "g1.cuh"

#define KZ_MSM_MAX_WINDOWS 96

struct MsmPlan {
    int nbits;                   // scalar width: 128 or 255
    int c;                       // window width
    int W;                       // number of windows = ceil(nbits / c)
    u32 nb[KZ_MSM_MAX_WINDOWS];          // buckets per window (top window is unsigned)
    u32 bucket_off[KZ_MSM_MAX_WINDOWS + 1];   // prefix sums of nb
    u32 total_buckets;
};

// Signed windows 0..W-2 (digit in [-2^(c-1), 2^(c-1)]), unsigned top window that absorbs the last carry,
// so no extra window is ever needed: top window holds tb = nbits - c(W-1) bits, values 0..2^tb.
inline MsmPlan msm_make_plan(size_t n, int nbits) {
    MsmPlan p;
    int lg = 0;
    while ((size_t(1) << (lg + 1)) <= n) ++lg;
    int c0 = lg - 3;
    if (c0 < 3) c0 = 3;
    if (c0 > 16) c0 = 16;
    // The top window holds tb = nbits - c(W-1) bits.  A narrow top window means a handful of giant buckets
    // (2^(c-1-tb) times the average load), which serialises the accumulation; pick, near c0, the cheapest
    // width whose top-window buckets are at most 8x heavier than the others.
    int c = c0;
    double best = -1;
    for (int cc = c0 - 3; cc <= c0 + 1; ++cc) {
        if (cc < 3 || cc > 16) continue;
        int W = (nbits + cc - 1) / cc, tb = nbits - cc * (W - 1);
        if (cc - 1 - tb > 3) continue;
        double cost = (double)W * (10.0 * (double)n + 28.0 * (double)(1u << (cc - 1)));
        if (best < 0 || cost < best) { best = cost; c = cc; }
    }
    p.nbits = nbits;
    p.c = c;
    p.W = (nbits + c - 1) / c;
    p.bucket_off[0] = 0;
    for (int w = 0; w < p.W; ++w) {
        int tb = nbits - c * (p.W - 1);
        p.nb[w] = (w < p.W - 1) ? (1u << (c - 1)) : (1u << tb);
        p.bucket_off[w + 1] = p.bucket_off[w] + p.nb[w];
    }
    p.total_buckets = p.bucket_off[p.W];
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

// GLV split of a canonical scalar k < r:  k = k1 + k2 * lambda  with  k2 = floor(k / lambda), 0 <= k1 < lambda,
// both below 2^128 (lambda = x^2 - 1, r = lambda^2 + lambda + 1).  phi(P) = (beta^2 x, y) = [lambda]P, so
// k*P = k1*P + k2*phi(P): a 255-bit sum over m points becomes a 128-bit sum over 2m points -- the same
// number of bucket additions, half the windows (bucket reductions, Horner doublings).
// Barrett quotient with mu = floor(2^384 / lambda): q^ = (k * mu) >> 384 is q or q-1.
KZ_HD void glv_split(const u32* k, u32* k1, u32* k2) {
    u32 prod[17];
    for (int i = 0; i < 17; ++i) prod[i] = 0;
    for (int i = 0; i < 8; ++i) {
        u64 c = 0;
        for (int j = 0; j < 9; ++j) {
            u64 v = (u64)k[i] * GLV_MU[j] + prod[i + j] + c;
            prod[i + j] = (u32)v;
            c = v >> 32;
        }
        prod[i + 9] = (u32)c;
    }
    u32 q[5];
    for (int i = 0; i < 5; ++i) q[i] = prod[12 + i];
    // rem = k - q * lambda  (fits in 5 limbs: < 2 lambda)
    u32 ql[9];
    for (int i = 0; i < 9; ++i) ql[i] = 0;
    for (int i = 0; i < 5; ++i) {
        u64 c = 0;
        for (int j = 0; j < 4; ++j) {
            u64 v = (u64)q[i] * GLV_LAMBDA[j] + ql[i + j] + c;
            ql[i + j] = (u32)v;
            c = v >> 32;
        }
        ql[i + 4] = (u32)c;
    }
    u32 rem[5];
    u64 bw = 0;
    for (int i = 0; i < 5; ++i) {
        u64 v = (u64)k[i] - ql[i] - bw;
        rem[i] = (u32)v;
        bw = (v >> 32) & 1;
    }
    for (int rep = 0; rep < 2; ++rep) {          // at most one correction is ever needed; two for safety
        u32 t[5];
        bw = 0;
        for (int i = 0; i < 5; ++i) {
            u64 v = (u64)rem[i] - (i < 4 ? GLV_LAMBDA[i] : 0u) - bw;
            t[i] = (u32)v;
            bw = (v >> 32) & 1;
        }
        if (!bw) {                                // rem >= lambda
            for (int i = 0; i < 5; ++i) rem[i] = t[i];
            u64 c = 1;
            for (int i = 0; i < 5; ++i) { u64 v = (u64)q[i] + c; q[i] = (u32)v; c = v >> 32; }
        }
    }
    for (int i = 0; i < 4; ++i) { k1[i] = rem[i]; k2[i] = q[i]; }
}
KZ_HD G1Aff g1_endo(const G1Aff& p) { return {fp_mul(fp_const(FP_BETA2), p.x), p.y}; }   // (0,0) stays (0,0)

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

// ---- balanced accumulation: every thread takes a fixed run of L sorted entries regardless of bucket
// boundaries, so a skewed bucket (e.g. the narrow top window) is spread over many threads.
//   pass 1: buckets lying entirely inside a chunk are written directly; the first / last run of a chunk
//           that continues into a neighbour chunk is parked as a partial record (head / tail).
//   pass 2: the chunk holding the first entry of such a bucket adds up the following head records.
#define KZ_KEY_NONE 0xFFFFFFFFu
struct ChunkRecs {
    G1Xyzz* head;        // [T] partial sum of the first run of chunk t
    G1Xyzz* tail;        // [T] partial sum of the last run of chunk t (when different from the first)
    u32* head_key;       // [T] bucket key or KZ_KEY_NONE
    u32* tail_key;       // [T]
    u32* head_flags;     // [T] bit0: run starts its bucket, bit1: run ends its bucket, bit2: run covers the whole chunk
};
KZ_HD void msm_chunk_pass1(const Fp* pts, const u32* keys, const u32* vals, u32 n_valid, u32 L, u32 t,
                           G1Xyzz* buckets, const ChunkRecs& R) {
    R.head_key[t] = KZ_KEY_NONE;
    R.tail_key[t] = KZ_KEY_NONE;
    u32 lo = t * L;
    if (lo >= n_valid) return;
    u32 hi = lo + L < n_valid ? lo + L : n_valid;
    u32 prev_key = lo ? keys[lo - 1] : KZ_KEY_NONE;
    u32 next_key = hi < n_valid ? keys[hi] : KZ_KEY_NONE;
    u32 cur = keys[lo];
    bool is_head = true;
    G1Xyzz acc = xyzz_inf();
    // software pipeline: the (key, value, point) of entry j+1 are loaded before entry j is added, so the
    // gather latency (a random 96-byte read per entry) hides under the ~11k-cycle mixed addition
    u32 k_next = cur, v_next = vals[lo];
    G1Aff p_next = load_point(pts, v_next & 0x7FFFFFFFu);
    for (u32 j = lo; j <= hi; ++j) {
        u32 k = j < hi ? k_next : KZ_KEY_NONE;
        u32 v = v_next;
        G1Aff p = p_next;
        if (j + 1 < hi) {
            k_next = keys[j + 1];
            v_next = vals[j + 1];
            p_next = load_point(pts, v_next & 0x7FFFFFFFu);
        }
        if (k != cur) {                                        // flush the finished run
            bool is_tail = j == hi;
            bool starts = is_head ? prev_key != cur : true;
            bool ends = is_tail ? next_key != cur : true;
            if (starts && ends) buckets[cur] = acc;
            else if (is_head) { R.head[t] = acc; R.head_key[t] = cur; R.head_flags[t] = (starts ? 1u : 0u) | (ends ? 2u : 0u) | (is_tail ? 4u : 0u); }
            else { R.tail[t] = acc; R.tail_key[t] = cur; }     // a tail that is not the head starts here and continues
            if (j == hi) break;
            cur = k;
            is_head = false;
            acc = xyzz_inf();
        }
        if (aff_is_inf(p)) continue;
        if (v >> 31) p.y = fp_neg(p.y);
        acc = xyzz_madd(acc, p);
    }
}
// The spanning bucket a chunk owns, if any.  Owner = the chunk where the bucket starts: its tail run, or a head run that
// starts the bucket and covers the whole chunk.  The two cases exclude each other (a head that covers the whole chunk
// is also the chunk's last run, so there is no separate tail), hence at most ONE owned bucket per chunk.
KZ_HD bool msm_chunk_owned(const ChunkRecs& R, u32 t, u32& key, G1Xyzz& sum) {
    key = R.tail_key[t];
    if (key != KZ_KEY_NONE) { sum = R.tail[t]; return true; }
    key = R.head_key[t];
    if (key == KZ_KEY_NONE) return false;
    const u32 f = R.head_flags[t];
    if (!(f & 1u) || (f & 2u)) return false;                  // does not start here, or already complete
    sum = R.head[t];
    return true;
}
KZ_HD void msm_chunk_pass2(u32 T, u32 t, G1Xyzz* buckets, const ChunkRecs& R) {
    u32 key;
    G1Xyzz sum;
    if (!msm_chunk_owned(R, t, key, sum)) return;
    for (u32 u = t + 1; u < T; ++u) {
        if (R.head_key[u] != key) break;
        sum = xyzz_add(sum, R.head[u]);
        if (R.head_flags[u] & 2u) break;
    }
    buckets[key] = sum;
}

// Horner over window sums: result = sum_w 2^(c w) * win[w].  The chain of c(W-1) doublings is serial, so it
// runs in Jacobian coordinates (2M+5S per doubling instead of 6M+3S in XYZZ).
KZ_COLD G1Jac msm_combine_body(const G1Xyzz* win, int W, int c) {
    G1Jac acc = jac_inf();
    for (int w = W - 1; w >= 0; --w) {
        if (!jac_is_inf(acc))
            for (int k = 0; k < c; ++k) acc = jac_dbl(acc);
        acc = jac_add(acc, xyzz_to_jac(win[w]));
    }
    return acc;
}

// ---- bucket reduction through row / column totals, and the batched subgroup check that shares them.
// A window with k magnitude bits (k = c-1, or tb for the unsigned top window) is a 2^kh x 2^kl table of buckets,
// m = row * 2^kl + col, kl = k / 2 (the bucket of magnitude m sits at index m - 1; m = 0 is no bucket, m = 2^k is
// the one bucket outside the table).  Bit b of m is a column bit (b < kl) or a row bit, so the slice sums
//     T_b = sum { B_m : bit b of m set }        b = 0 .. k-1
// are sums of column totals Q_c or of row totals R_r, and the window total is a Horner chain over them:
//     sum_m m B_m = 2^k B_{2^k} + sum_b 2^b T_b.
//   pass 1  sums of runs of KZ_SG_RUN buckets along rows and along columns (every lane busy)
//   pass 2  the 2^kh row totals and 2^kl column totals
//   slices  T_b from the totals (one block per slice), plus, for the batched subgroup check, the slice "all" =
//           sum of every bucket of a signed window
//   window  Horner over the k slices of each window (one thread per window)
// 2 * 2^k additions per window, all of them independent, then 2k serial point operations.
//
// Batched subgroup check: the signed digits of the challenges r_i are 128 fair, independent coins per
// point -- for a signed window the c-1 magnitude bits and the sign, for the top window its tb bits.  Point P_i
// enters slice (w, b) with coefficient sign * bit in {-1, 0, 1} and slice (w, all) with its sign; each value has
// probability <= 1/2, independently over the 128 slices, so a point with a non-zero cofactor component (odd
// order >= 3) leaves all 128 slice sums inside G1 with probability <= 2^-128.  Checking the 128 sums of
// sum r_i C_i and of sum r_i pi_i replaces 2n per-point checks, with no extra point additions.
#define KZ_SG_RUN 16
struct SgWin { int k, kl, kh; u32 rows, cols, tpr, tpc, rjobs, jobs; };
KZ_HD SgWin sg_win(const MsmPlan& P, int w) {
    SgWin g;
    g.k = w < P.W - 1 ? P.c - 1 : P.nbits - P.c * (P.W - 1);
    g.kl = g.k / 2; g.kh = g.k - g.kl;
    g.rows = 1u << g.kh; g.cols = 1u << g.kl;
    g.tpr = g.cols > KZ_SG_RUN ? g.cols / KZ_SG_RUN : 1u;       // threads per row
    g.tpc = g.rows > KZ_SG_RUN ? g.rows / KZ_SG_RUN : 1u;       // threads per column
    g.rjobs = g.rows * g.tpr;
    g.jobs = g.rjobs + g.cols * g.tpc;
    return g;
}
// scratch layout of one sum: [W][stride] run sums, then [W][tstride] totals (R then Q); strides of the widest window
struct SgLayout { u32 stride, tstride; };
inline SgLayout sg_layout(const MsmPlan& P) {
    int kmax = P.c - 1, tb = P.nbits - P.c * (P.W - 1);
    if (P.W == 1 || tb > kmax) kmax = tb;
    int kl = kmax / 2, kh = kmax - kl;              // floor and ceil are monotone in k: the widest window bounds all
    u32 rows = 1u << kh, cols = 1u << kl;
    u32 tpr = cols > KZ_SG_RUN ? cols / KZ_SG_RUN : 1u, tpc = rows > KZ_SG_RUN ? rows / KZ_SG_RUN : 1u;
    return {rows * tpr + cols * tpc, rows + cols};
}
inline size_t sg_work_entries(const MsmPlan& P) {
    SgLayout L = sg_layout(P);
    return (size_t)P.W * (L.stride + L.tstride);
}
KZ_HD G1Xyzz sg_bucket(const G1Xyzz* base, u32 m) { return m ? base[m - 1] : xyzz_inf(); }
// pass 1, job t of window g: a run along a row (t < rjobs) or along a column
KZ_HD G1Xyzz sg_run_sum(const G1Xyzz* base, const SgWin& g, u32 t) {
    G1Xyzz acc = xyzz_inf();
    if (t < g.rjobs) {
        u32 r = t / g.tpr, p = t - r * g.tpr, len = g.cols / g.tpr;
        u32 m0 = (r << g.kl) + p * len;
        for (u32 i = 0; i < len; ++i) acc = xyzz_add(acc, sg_bucket(base, m0 + i));
    } else {
        u32 u = t - g.rjobs, c = u / g.tpc, p = u - c * g.tpc, len = g.rows / g.tpc;
        for (u32 i = 0; i < len; ++i) acc = xyzz_add(acc, sg_bucket(base, ((p * len + i) << g.kl) + c));
    }
    return acc;
}
// pass 2, job t: row total R[t] (t < rows) or column total Q[t - rows] from the run sums of the window
KZ_HD G1Xyzz sg_total(const G1Xyzz* part_w, const SgWin& g, u32 t) {
    const G1Xyzz* src = part_w + (t < g.rows ? t * g.tpr : g.rjobs + (t - g.rows) * g.tpc);
    const u32 cnt = t < g.rows ? g.tpr : g.tpc;
    G1Xyzz acc = src[0];
    for (u32 i = 1; i < cnt; ++i) acc = xyzz_add(acc, src[i]);
    return acc;
}
// slice ids: signed window w: w*c + b, b < c-1 magnitude bits, b = c-1 the slice "all"; top window: c(W-1) + b
struct SgSlice { int w; int b; bool all; };
KZ_HD SgSlice sg_slice(const MsmPlan& P, int sid) {
    SgSlice s;
    int signed_slices = P.c * (P.W - 1);
    if (sid < signed_slices) { s.w = sid / P.c; s.b = sid - s.w * P.c; s.all = s.b == P.c - 1; }
    else { s.w = P.W - 1; s.b = sid - signed_slices; s.all = false; }
    return s;
}
// the part of a slice sum that lane `lane` of `nl` lanes adds up (tot_w = R then Q of the window)
KZ_HD G1Xyzz sg_slice_part(const G1Xyzz* tot_w, const G1Xyzz* base, const SgWin& g, const SgSlice& sl, u32 lane, u32 nl) {
    const G1Xyzz* R = tot_w;
    const G1Xyzz* Q = tot_w + g.rows;
    G1Xyzz acc = xyzz_inf();
    if (sl.all) {
        for (u32 r = lane; r < g.rows; r += nl) acc = xyzz_add(acc, R[r]);
        if (lane == 0) acc = xyzz_add(acc, base[(1u << g.k) - 1u]);            // magnitude 2^k
    } else if (sl.b < g.kl) {
        for (u32 c = lane; c < g.cols; c += nl) if ((c >> sl.b) & 1u) acc = xyzz_add(acc, Q[c]);
    } else {
        for (u32 r = lane; r < g.rows; r += nl) if ((r >> (sl.b - g.kl)) & 1u) acc = xyzz_add(acc, R[r]);
    }
    return acc;
}
// window total from its k bit slices T[0..k) and the bucket of magnitude 2^k
KZ_COLD G1Xyzz msm_window_from_slices(const G1Xyzz* T, int k, const G1Xyzz& ext) {
    G1Xyzz acc = ext;
    for (int b = k - 1; b >= 0; --b) acc = xyzz_add(xyzz_dbl(acc), T[b]);
    return acc;
}
KZ_COLD bool sg_sum_in_g1(const G1Xyzz& t) {
    if (xyzz_is_inf(t)) return true;
    G1Jac j = xyzz_to_jac(t);                    // no inversion on the serial path: both chains and the test are projective
    return g1_subgroup_compare_jac(j, jac_mul_xabs(jac_mul_xabs(j)));
}
