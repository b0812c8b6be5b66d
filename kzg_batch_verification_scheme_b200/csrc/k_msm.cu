// K4-K7: Pippenger MSM kernels (BASELINE.json:5 item (d)).
//   K4 digits     one thread per scalar: signed-window recoding -> (bucket key, point|sign) pairs
//   K5 sort       hand-written counting sort of the pairs by bucket key: the key space is the bucket table itself
//                 (<= 2^19 keys), so one histogram pass (fused into K4), one exclusive scan -- which IS the bucket
//                 boundary table -- and one scatter pass replace a multi-pass radix sort.  The order inside a
//                 bucket is arbitrary; bucket sums are group elements, so every canonical output is unchanged.
//   K6 accumulate bucket sums in XYZZ (mixed additions with the affine points)
//   K7 reduce     per-segment running sums, per-window block reduction, Horner combine over windows
#include <cstdlib>

#include "kernels.h"
#include "quad.cuh"

__global__ void __launch_bounds__(128) k_msm_digits(const u32* __restrict__ scalars, int nl, size_t m,
                                                    u32* __restrict__ keys, u32* __restrict__ vals, u32* __restrict__ count,
                                                    const __grid_constant__ MsmPlan plan) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u32 sc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = k < nl ? scalars[(size_t)nl * i + k] : 0u;
    msm_digits_body(keys, vals, sc, i, m, plan);
    // histogram of the keys just written (total_buckets + 1 counters; the last one collects zero digits)
    for (int w = 0; w < plan.W; ++w) atomicAdd(&count[keys[(size_t)w * m + i]], 1u);
}

// Exclusive scan of count[0 .. nkeys) into start[0 .. nkeys] (start[nkeys] = N) and a working copy `cursor`: the bucket
// boundary table.  Three small launches of 256-thread blocks (tile sums, scan of the <= 1024 tile sums, tile scans):
// blocks this small fit beside the resident K1 blocks, so the side stream does not wait for a K1 wave to retire -- the
// single 1024-thread block of round 1 did (and took 0.44 ms per sum at n = 2^20 on one SM).
#define KZ_SCAN_THREADS 256
#define KZ_SCAN_ITEMS 8
#define KZ_SCAN_TILE (KZ_SCAN_THREADS * KZ_SCAN_ITEMS)
#define KZ_SCAN_MAX_TILES 1024
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* sh, u32& total) {
    // Hillis-Steele over KZ_SCAN_THREADS values in shared memory; returns the exclusive prefix of this thread
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < KZ_SCAN_THREADS; off <<= 1) {
        u32 t = (int)threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    total = sh[KZ_SCAN_THREADS - 1];
    u32 incl = sh[threadIdx.x];
    __syncthreads();
    return incl - v;
}
__global__ void __launch_bounds__(KZ_SCAN_THREADS) k_scan_tiles(const u32* __restrict__ count, u32 nkeys, u32 tiles_per_block,
                                                                u32* __restrict__ tile_sum) {
    __shared__ u32 sh[KZ_SCAN_THREADS];
    // block b sums tiles_per_block consecutive tiles (so that at most KZ_SCAN_MAX_TILES sums remain)
    const u32 lo = blockIdx.x * tiles_per_block * KZ_SCAN_TILE;
    u32 hi = lo + tiles_per_block * KZ_SCAN_TILE;
    if (hi > nkeys) hi = nkeys;
    u32 sum = 0;
    for (u32 k = lo + threadIdx.x; k < hi; k += KZ_SCAN_THREADS) sum += count[k];
    u32 total;
    block_exclusive_scan(sum, sh, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(KZ_SCAN_THREADS) k_scan_tile_sums(u32* __restrict__ tile_sum, u32 nblocks, u32 nkeys, u32* __restrict__ start) {
    __shared__ u32 sh[KZ_SCAN_THREADS];
    // <= KZ_SCAN_MAX_TILES values: each thread owns a contiguous slice of 4
    const u32 per = (nblocks + KZ_SCAN_THREADS - 1) / KZ_SCAN_THREADS;
    const u32 lo = threadIdx.x * per, hi = lo + per < nblocks ? lo + per : nblocks;
    u32 sum = 0;
    for (u32 k = lo; k < hi; ++k) sum += tile_sum[k];
    u32 total;
    u32 run = block_exclusive_scan(sum, sh, total);
    for (u32 k = lo; k < hi; ++k) { u32 v = tile_sum[k]; tile_sum[k] = run; run += v; }
    if (threadIdx.x == 0) start[nkeys] = total;
}
__global__ void __launch_bounds__(KZ_SCAN_THREADS) k_scan_final(const u32* __restrict__ count, u32 nkeys, u32 tiles_per_block,
                                                                const u32* __restrict__ tile_off, u32* __restrict__ start,
                                                                u32* __restrict__ cursor) {
    __shared__ u32 sh[KZ_SCAN_THREADS];
    u32 base = tile_off[blockIdx.x];
    for (u32 t = 0; t < tiles_per_block; ++t) {
        const u32 lo = (blockIdx.x * tiles_per_block + t) * KZ_SCAN_TILE + threadIdx.x * KZ_SCAN_ITEMS;
        u32 v[KZ_SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int i = 0; i < KZ_SCAN_ITEMS; ++i) { v[i] = lo + i < nkeys ? count[lo + i] : 0u; sum += v[i]; }
        u32 total;
        u32 run = base + block_exclusive_scan(sum, sh, total);
#pragma unroll
        for (int i = 0; i < KZ_SCAN_ITEMS; ++i) {
            if (lo + i < nkeys) { start[lo + i] = run; cursor[lo + i] = run; }
            run += v[i];
        }
        base += total;
    }
}
static void launch_bucket_scan(cudaStream_t s, const u32* count, u32 nkeys, u32* start, u32* cursor, u32* tile_sum) {
    const u32 tiles = (nkeys + KZ_SCAN_TILE - 1) / KZ_SCAN_TILE;
    const u32 tpb = (tiles + KZ_SCAN_MAX_TILES - 1) / KZ_SCAN_MAX_TILES;          // tiles per block: 1 up to 2 M keys
    const u32 nblocks = (tiles + tpb - 1) / tpb;
    k_scan_tiles<<<nblocks, KZ_SCAN_THREADS, 0, s>>>(count, nkeys, tpb, tile_sum);
    KZ_COUNT_LAUNCH();
    k_scan_tile_sums<<<1, KZ_SCAN_THREADS, 0, s>>>(tile_sum, nblocks, nkeys, start);
    KZ_COUNT_LAUNCH();
    k_scan_final<<<nblocks, KZ_SCAN_THREADS, 0, s>>>(count, nkeys, tpb, tile_sum, start, cursor);
    KZ_COUNT_LAUNCH();
}
__global__ void __launch_bounds__(256) k_bucket_scatter(const u32* __restrict__ keys, const u32* __restrict__ vals, size_t N,
                                                        u32* __restrict__ cursor, u32* __restrict__ skeys, u32* __restrict__ svals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    u32 k = keys[i];
    u32 pos = atomicAdd(&cursor[k], 1u);
    skeys[pos] = k;
    svals[pos] = vals[i];
}

// GLV preparation of a 255-bit sum: scalars k -> (k1 | k2) as 2m 128-bit scalars, points P -> phi(P)
__global__ void __launch_bounds__(128) k_glv_split(const u32* __restrict__ scalars8, size_t m, u32* __restrict__ out4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u32 sc[8], k1[4], k2[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = scalars8[8 * i + j];
    glv_split(sc, k1, k2);
    reinterpret_cast<uint4*>(out4)[i] = make_uint4(k1[0], k1[1], k1[2], k1[3]);
    reinterpret_cast<uint4*>(out4)[m + i] = make_uint4(k2[0], k2[1], k2[2], k2[3]);
}
__global__ void __launch_bounds__(128) k_endo_points(const Fp* __restrict__ src, size_t m, Fp* __restrict__ dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    G1Aff q = g1_endo(load_point(src, i));
    dst[2 * i] = q.x;
    dst[2 * i + 1] = q.y;
}
void launch_glv_split(cudaStream_t s, const uint32_t* scalars8, size_t m, uint32_t* out4) {
    if (!m) return;
    k_glv_split<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(scalars8, m, out4);
    KZ_COUNT_LAUNCH();
}
void launch_endo_points(cudaStream_t s, const Fp* src, size_t m, Fp* dst) {
    if (!m) return;
    k_endo_points<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(src, m, dst);
    KZ_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(128) k_msm_chunk_pass1(const Fp* __restrict__ pts, const u32* __restrict__ keys,
                                                         const u32* __restrict__ vals, const u32* __restrict__ start,
                                                         u32 total_buckets, u32 L, u32 T, G1Xyzz* __restrict__ buckets,
                                                         ChunkRecs R) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    msm_chunk_pass1(pts, keys, vals, start[total_buckets], L, t, buckets, R);
}
// Warp-cooperative variant of pass 1 (BASELINE.json:5 item (d): "points staged through shared memory and read
// with coalesced 128-bit loads").  All 32 lanes walk their runs in lock step; for every step the warp gathers
// its 32 affine points together: lane l loads 16-byte piece (q*32 + l) % 6 of point (q*32 + l) / 6, so six
// consecutive lanes read one contiguous 96-byte point and every load instruction touches fully used sectors
// (the per-lane version reads each 32-byte sector twice).  The pieces go to a per-warp shared-memory tile and
// each lane reads its own point back with six 128-bit shared loads.  The gather for step s+1 is issued
// before the mixed addition of step s (registers), written to the tile afterwards.
#define KZ_ACC_WARPS 4
template <int WARPS, int MAXREG>
__global__ void __launch_bounds__(32 * WARPS) __maxnreg__(MAXREG) k_msm_chunk_pass1_staged(const Fp* __restrict__ pts, const u32* __restrict__ keys,
                                                                              const u32* __restrict__ vals, const u32* __restrict__ start,
                                                                              u32 total_buckets, u32 L, u32 T,
                                                                              G1Xyzz* __restrict__ buckets, ChunkRecs R) {
    __shared__ uint4 tile[WARPS][32 * 6];
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint4* my_tile = tile[wid];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n_valid = start[total_buckets];
    const bool active = t < T && t * L < n_valid;
    if (t < T) { R.head_key[t] = KZ_KEY_NONE; R.tail_key[t] = KZ_KEY_NONE; }
    const u32 lo = t * L;
    const u32 hi = active ? (lo + L < n_valid ? lo + L : n_valid) : lo;
    const u32 prev_key = (active && lo) ? keys[lo - 1] : KZ_KEY_NONE;
    const u32 next_key = (active && hi < n_valid) ? keys[hi] : KZ_KEY_NONE;
    u32 cur = active ? keys[lo] : KZ_KEY_NONE;
    bool is_head = true, done = !active;
    G1Xyzz acc = xyzz_inf();
    const uint4* base = reinterpret_cast<const uint4*>(pts);
    // gather of one step into registers: 6 pieces per lane, point index fetched from the owning lane
    uint4 piece[6];
    auto gather = [&](u32 v_mine, bool have_mine) {
        u32 idx_mine = have_mine ? (v_mine & 0x7FFFFFFFu) : 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            u32 g = (u32)q * 32u + lane, owner = g / 6u, part = g - owner * 6u;
            u32 idx = __shfl_sync(0xFFFFFFFFu, idx_mine, owner);
            piece[q] = idx != 0xFFFFFFFFu ? __ldg(base + (size_t)idx * 6 + part) : make_uint4(0, 0, 0, 0);
        }
    };
    auto publish = [&]() {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 6; ++q) my_tile[q * 32 + lane] = piece[q];
        __syncwarp();
    };
    u32 v_next = active ? vals[lo] : 0u;
    gather(v_next, active);
    publish();
    for (u32 step = 0; step <= L; ++step) {                    // warp-uniform trip count
        const u32 j = lo + step;
        const bool have = !done && j < hi;
        const u32 k = have ? keys[j] : KZ_KEY_NONE;
        const u32 v = v_next;
        // my point of this step from the tile
        G1Aff p;
        {
            const uint4* mine = my_tile + lane * 6;
            uint4 a = mine[0], b = mine[1], c = mine[2], d = mine[3], e = mine[4], f = mine[5];
            p.x.v[0] = a.x; p.x.v[1] = a.y; p.x.v[2] = a.z; p.x.v[3] = a.w; p.x.v[4] = b.x; p.x.v[5] = b.y; p.x.v[6] = b.z; p.x.v[7] = b.w;
            p.x.v[8] = c.x; p.x.v[9] = c.y; p.x.v[10] = c.z; p.x.v[11] = c.w;
            p.y.v[0] = d.x; p.y.v[1] = d.y; p.y.v[2] = d.z; p.y.v[3] = d.w; p.y.v[4] = e.x; p.y.v[5] = e.y; p.y.v[6] = e.z; p.y.v[7] = e.w;
            p.y.v[8] = f.x; p.y.v[9] = f.y; p.y.v[10] = f.z; p.y.v[11] = f.w;
        }
        // issue the gather of the next step before the arithmetic of this one
        const bool have_next = !done && j + 1 < hi;
        v_next = have_next ? vals[j + 1] : 0u;
        if (step < L) gather(v_next, have_next);
        if (!done && k != cur) {                               // flush the finished run (k == NONE at j == hi)
            const bool is_tail = j == hi;
            const bool starts = is_head ? prev_key != cur : true;
            const bool ends = is_tail ? next_key != cur : true;
            if (starts && ends) buckets[cur] = acc;
            else if (is_head) { R.head[t] = acc; R.head_key[t] = cur; R.head_flags[t] = (starts ? 1u : 0u) | (ends ? 2u : 0u) | (is_tail ? 4u : 0u); }
            else { R.tail[t] = acc; R.tail_key[t] = cur; }
            if (is_tail) done = true;
            cur = k;
            is_head = false;
            acc = xyzz_inf();
        }
        if (have && !aff_is_inf(p)) {
            if (v >> 31) p.y = fp_neg(p.y);
            acc = xyzz_madd(acc, p);
        }
        if (step < L) publish();
    }
}

__global__ void __launch_bounds__(128) k_msm_chunk_pass2_t(u32 T, G1Xyzz* __restrict__ buckets, ChunkRecs R) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    msm_chunk_pass2(T, t, buckets, R);
}
// Pass 2, one QUAD per chunk: the partial records of a bucket that straddles chunks are added up by the
// chunk where the bucket starts (msm_chunk_pass2 in msm.cuh is the one-thread form the emulation runs).  The additions
// are a serial chain per bucket, so small sums run them on quads (n = 4096: accumulate stage 0.45 -> 0.37 ms); with many
// chunks the one-thread form has the better throughput (n = 2^20: 15.2 vs 16.3 ms), see msm_accumulate_stage.
__global__ void __launch_bounds__(128) k_msm_chunk_pass2(u32 T, G1Xyzz* __restrict__ buckets, ChunkRecs R) {
    __shared__ Fp qsm[32 * KZ_QUAD_SLOTS];
    const int qi = threadIdx.x >> 2;
    const u32 t = blockIdx.x * 32u + (u32)qi;
    if (t >= T) return;
    u32 key;
    G1Xyzz sum;
    if (!msm_chunk_owned(R, t, key, sum)) return;
    Quad q = quad_make(qsm, qi);
    for (u32 u = t + 1; u < T; ++u) {
        if (R.head_key[u] != key) break;
        sum = quad_xyzz_add(q, sum, R.head[u]);
        if (R.head_flags[u] & 2u) break;
    }
    if (q.ql == 0) buckets[key] = sum;
}

// ---- bucket reduction through row / column totals (msm.cuh) with quad-lane point arithmetic (quad.cuh).  The run-sum
// kernels of round 1 (one thread per run of 16 buckets) were measured no faster even at 2^15 buckets per window
// (n = 2^20: 46.10 vs 46.14 ms) and 3x slower on small tables, and were dropped.
// One block of KZ_RED_THREADS / 4 quads per row / column total: strided partial sums, then a tree over the quads.
#define KZ_RED_THREADS 64
__global__ void __launch_bounds__(KZ_RED_THREADS) k_red_totals(const G1Xyzz* __restrict__ buckets, G1Xyzz* __restrict__ tot, u32 tstride,
                                                               const __grid_constant__ MsmPlan plan) {
    constexpr int NQ = KZ_RED_THREADS / 4;
    __shared__ Fp qsm[NQ * KZ_QUAD_SLOTS];
    __shared__ G1Xyzz red[NQ];
    const int w = blockIdx.y;
    const SgWin g = sg_win(plan, w);
    const u32 t = blockIdx.x;
    if (t >= g.rows + g.cols) return;                       // uniform over the block
    const int qi = threadIdx.x >> 2;
    Quad q = quad_make(qsm, qi);
    const G1Xyzz* base = buckets + plan.bucket_off[w];
    G1Xyzz acc = xyzz_inf();
    if (t < g.rows) {
        for (u32 e = qi; e < g.cols; e += NQ) acc = quad_xyzz_add(q, acc, sg_bucket(base, (t << g.kl) + e));
    } else {
        const u32 c = t - g.rows;
        for (u32 e = qi; e < g.rows; e += NQ) acc = quad_xyzz_add(q, acc, sg_bucket(base, (e << g.kl) + c));
    }
    for (int st = NQ / 2; st > 0; st >>= 1) {               // level st: quads [st, 2st) publish, quads [0, st) add
        if (qi >= st && qi < 2 * st && q.ql == 0) red[qi] = acc;
        __syncthreads();
        if (qi < st) acc = quad_xyzz_add(q, acc, red[qi + st]);
    }
    if (threadIdx.x == 0) tot[(size_t)w * tstride + t] = acc;
}
// one block of 32 quads per slice sum
#define KZ_SLICE_THREADS 128
__global__ void __launch_bounds__(KZ_SLICE_THREADS) k_red_slices(const G1Xyzz* __restrict__ buckets, const G1Xyzz* __restrict__ tot, u32 tstride,
                                                                 G1Xyzz* __restrict__ slices, int want_all, const __grid_constant__ MsmPlan plan) {
    constexpr int NQ = KZ_SLICE_THREADS / 4;
    __shared__ Fp qsm[NQ * KZ_QUAD_SLOTS];
    __shared__ G1Xyzz red[NQ];
    const SgSlice sl = sg_slice(plan, (int)blockIdx.x);
    if (sl.all && !want_all) return;
    const SgWin g = sg_win(plan, sl.w);
    const int qi = threadIdx.x >> 2;
    Quad q = quad_make(qsm, qi);
    const G1Xyzz* R = tot + (size_t)sl.w * tstride;
    const G1Xyzz* Q = R + g.rows;
    G1Xyzz acc = xyzz_inf();
    if (sl.all) {
        for (u32 r = qi; r < g.rows; r += NQ) acc = quad_xyzz_add(q, acc, R[r]);
        if (qi == 0) acc = quad_xyzz_add(q, acc, buckets[plan.bucket_off[sl.w] + (1u << g.k) - 1u]);       // magnitude 2^k
    } else if (sl.b < g.kl) {
        // the columns with bit b set, enumerated densely: c = (hi << (b+1)) | 1 << b | lo
        const u32 half = g.cols >> 1, lowmask = (1u << sl.b) - 1u;
        for (u32 i = qi; i < half; i += NQ) acc = quad_xyzz_add(q, acc, Q[((i & ~lowmask) << 1) | (1u << sl.b) | (i & lowmask)]);
    } else {
        const u32 b = sl.b - g.kl, half = g.rows >> 1, lowmask = (1u << b) - 1u;
        for (u32 i = qi; i < half; i += NQ) acc = quad_xyzz_add(q, acc, R[((i & ~lowmask) << 1) | (1u << b) | (i & lowmask)]);
    }
    for (int st = NQ / 2; st > 0; st >>= 1) {
        if (qi >= st && qi < 2 * st && q.ql == 0) red[qi] = acc;
        __syncthreads();
        if (qi < st) acc = quad_xyzz_add(q, acc, red[qi + st]);
    }
    if (threadIdx.x == 0) slices[blockIdx.x] = acc;
}
// one quad per window: Horner over its bit slices
__global__ void __launch_bounds__(32) k_msm_window_horner_q(const G1Xyzz* __restrict__ buckets, const G1Xyzz* __restrict__ slices,
                                                             G1Xyzz* __restrict__ winsums, const __grid_constant__ MsmPlan plan) {
    __shared__ Fp qsm[8 * KZ_QUAD_SLOTS];
    const int qi = threadIdx.x >> 2, w = blockIdx.x * 8 + qi;
    if (w >= plan.W) return;
    Quad q = quad_make(qsm, qi);
    const SgWin g = sg_win(plan, w);
    const G1Xyzz* T = slices + (size_t)w * plan.c;
    G1Xyzz acc = buckets[plan.bucket_off[w] + (1u << g.k) - 1u];
    for (int b = g.k - 1; b >= 0; --b) acc = quad_xyzz_add(q, quad_xyzz_dbl(q, acc), T[b]);
    if (q.ql == 0) winsums[w] = acc;
}
// batched subgroup check: one quad per slice sum of the two sums; counters[2] += sums outside G1
__global__ void __launch_bounds__(32) k_sg_check_q(const G1Xyzz* __restrict__ slices_a, const G1Xyzz* __restrict__ slices_b, int nslices,
                                                    u32* __restrict__ counters) {
    __shared__ Fp qsm[8 * KZ_QUAD_SLOTS];
    const int qi = threadIdx.x >> 2, sid = blockIdx.x * 8 + qi;
    if (sid >= nslices) return;
    Quad q = quad_make(qsm, qi);
    const bool ok = quad_sum_in_g1(q, (blockIdx.y ? slices_b : slices_a)[sid]);
    if (!ok && q.ql == 0) atomicAdd(counters + 2, 1u);
}
// Horner combine; one block per job so the three sums of a batch run their serial chains concurrently
struct CombineJobs {
    const G1Xyzz* winsums[3];
    G1Jac* out[3];
    int W[3], c[3];
};
__global__ void __launch_bounds__(32) k_msm_combine(CombineJobs jobs) {
    __shared__ Fp qsm[KZ_QUAD_SLOTS];
    if (threadIdx.x >= 4) return;
    Quad q = quad_make(qsm, 0);
    const int j = blockIdx.x;
    const G1Xyzz* win = jobs.winsums[j];
    G1Xyzz acc = xyzz_inf();
    for (int w = jobs.W[j] - 1; w >= 0; --w) {
        if (!xyzz_is_inf(acc))
            for (int k = 0; k < jobs.c[j]; ++k) acc = quad_xyzz_dbl(q, acc);
        acc = quad_xyzz_add(q, acc, win[w]);
    }
    if (q.ql == 0) *jobs.out[j] = xyzz_to_jac(acc);
}

void msm_sort_stage(cudaStream_t s, const MsmPlan& plan, const uint32_t* scalars, int nl, size_t m, MsmWorkspace& ws) {
    size_t N = m * (size_t)plan.W;
    u32 nkeys = plan.total_buckets + 1;                  // + the "zero digit" key
    cudaMemsetAsync(ws.count, 0, sizeof(u32) * (nkeys + 1), s);
    k_msm_digits<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(scalars, nl, m, ws.keys_alt, ws.vals_alt, ws.count, plan);
    KZ_COUNT_LAUNCH();
    launch_bucket_scan(s, ws.count, nkeys, ws.bucket_start, ws.cursor, ws.tile_sum);
    k_bucket_scatter<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(ws.keys_alt, ws.vals_alt, N, ws.cursor, ws.keys, ws.vals);
    KZ_COUNT_LAUNCH();
}
u32 msm_chunk_len(size_t N) { return N >= (1u << 21) ? 32u : N >= (1u << 20) ? 16u : N >= (1u << 18) ? 8u : 4u; }
// Chunk length of one sum.  Pass 2 adds the partial records of a bucket one after the other, load / L general
// additions for the heaviest buckets (the narrow top window), while pass 1 runs L mixed additions per thread; for
// small sums both are serial latency, balanced at L ~ sqrt(1.4 * heaviest expected load).  Large sums are
// throughput-bound and keep the length chosen from the entry count.
static u32 msm_chunk_len_plan(const MsmPlan& plan, size_t m) {
    u32 base = msm_chunk_len(m * (size_t)plan.W);
    u32 nb_min = plan.nb[0];
    for (int w = 1; w < plan.W; ++w) if (plan.nb[w] < nb_min) nb_min = plan.nb[w];
    double heavy = (double)m / (double)(nb_min ? nb_min : 1);
    u32 L = 4;
    while (L < 32 && (double)L * (double)L < 1.4 * heavy) L <<= 1;
    L = L > base ? L : base;
    // Wave quantisation: pass 1 runs ceil(T / 128) blocks, `per_wave` of them at a time (2 per SM at 232 registers), every
    // block for L mixed additions -- 2.16 waves cost as much as 3.  For sums of up to a few waves pick, in [L, 2L), the
    // length with the fewest (waves x L) (n = 2^17: 640 blocks of 16 entries = 3 x 16 -> 569 blocks of 18 = 2 x 18).
    static const u32 per_wave = [] {
        int dev = 0, sms = 148, occ = 2;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_msm_chunk_pass1_staged<KZ_ACC_WARPS, 232>, 32 * KZ_ACC_WARPS, 0);
        return (u32)(sms * (occ > 0 ? occ : 1));
    }();
    static const int tune = [] { const char* e = getenv("KZGB_ACC_WAVE_TUNE"); return e ? atoi(e) : 1; }();
    const size_t N = m * (size_t)plan.W;
    // Very large sums (the window width is capped at 16, so buckets hold hundreds of entries from m = 2^22 on): with L = 32 a
    // bucket straddles dozens of chunks and pass 2 adds their records one after the other on ONE lane (m = 2^24, 255-bit:
    // 238 ms instead of ~110).  Grow the chunk until a bucket spans ~4 chunks, keeping at least 8 waves of pass-1 threads.
    while (L < 1024 && (double)L < heavy / 4.0 && N / (2 * (size_t)L) >= 8 * (size_t)per_wave * 32 * KZ_ACC_WARPS) L <<= 1;
    if (tune && N / L / 128 < 12 * (size_t)per_wave) {
        u32 bestL = L;
        double best = 1e300;
        for (u32 l = L; l < 2 * L; ++l) {
            const size_t blocks = ((N + l - 1) / l + 127) / 128;
            const double cost = (double)((blocks + per_wave - 1) / per_wave) * (double)l;
            if (cost < best - 1e-9) { best = cost; bestL = l; }
        }
        L = bestL;
    }
    return L;
}

void msm_accumulate_stage(cudaStream_t s, const MsmPlan& plan, const Fp* pts, size_t m, MsmWorkspace& ws) {
    size_t N = m * (size_t)plan.W;
    u32 L = msm_chunk_len_plan(plan, m);
    u32 T = (u32)((N + L - 1) / L);
    cudaMemsetAsync(ws.buckets, 0, sizeof(G1Xyzz) * (size_t)plan.total_buckets, s);      // empty buckets = infinity
    static const int staged = [] { const char* e = getenv("KZGB_ACC_STAGED"); return e ? atoi(e) : 1; }();
    if (staged) {
        // register budget and block shape measured on B200 (n = 2^20): 4 warps x 232 registers = 16.9 ms for the
        // three sums; 168 registers (3 blocks/SM) 18.4 ms, 1-2 warp blocks at 184-255 registers 17.0-18.6 ms
        k_msm_chunk_pass1_staged<KZ_ACC_WARPS, 232><<<(T + 32 * KZ_ACC_WARPS - 1) / (32 * KZ_ACC_WARPS), 32 * KZ_ACC_WARPS, 0, s>>>(
            pts, ws.keys, ws.vals, ws.bucket_start, plan.total_buckets, L, T, ws.buckets, ws.recs);
    } else {
        k_msm_chunk_pass1<<<(T + 127) / 128, 128, 0, s>>>(pts, ws.keys, ws.vals, ws.bucket_start, plan.total_buckets, L, T,
                                                          ws.buckets, ws.recs);
    }
    KZ_COUNT_LAUNCH();
    if (T <= 16384) k_msm_chunk_pass2<<<(T + 31) / 32, 128, 0, s>>>(T, ws.buckets, ws.recs);
    else k_msm_chunk_pass2_t<<<(T + 127) / 128, 128, 0, s>>>(T, ws.buckets, ws.recs);
    KZ_COUNT_LAUNCH();
}
// Slice sums of one sum: ws.slices[0 .. nbits) (sg_slice numbering).  want_all: also the "all" slices of the signed
// windows (needed only by the batched subgroup check).
void msm_slices_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws, bool want_all) {
    const SgLayout L = sg_layout(plan);
    G1Xyzz* tot = ws.sg_work + (size_t)plan.W * L.stride;
    k_red_totals<<<dim3(L.tstride, (unsigned)plan.W), KZ_RED_THREADS, 0, s>>>(ws.buckets, tot, L.tstride, plan);
    KZ_COUNT_LAUNCH();
    k_red_slices<<<(unsigned)plan.nbits, KZ_SLICE_THREADS, 0, s>>>(ws.buckets, tot, L.tstride, ws.slices, want_all ? 1 : 0, plan);
    KZ_COUNT_LAUNCH();
}
// per-window totals (ws.winsums) from the slice sums: only the classic path (Horner combine) needs them
void msm_winsums_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws) {
    k_msm_window_horner_q<<<(unsigned)((plan.W + 7) / 8), 32, 0, s>>>(ws.buckets, ws.slices, ws.winsums, plan);
    KZ_COUNT_LAUNCH();
}
void msm_window_sums_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws, bool want_all) {
    msm_slices_stage(s, plan, ws, want_all);
    msm_winsums_stage(s, plan, ws);
}
// batched subgroup check on the slice sums msm_slices_stage(.., want_all = true) left in the two workspaces
// (same plan): the 2 x 128 |x|^2 chains run side by side in one launch, one quad each
void launch_sg_check(cudaStream_t s, const MsmPlan& plan, const MsmWorkspace& wa, const MsmWorkspace& wb, uint32_t* counters, int nsums) {
    k_sg_check_q<<<dim3((unsigned)((plan.nbits + 7) / 8), (unsigned)nsums), 32, 0, s>>>(wa.slices, wb.slices, plan.nbits, counters);
    KZ_COUNT_LAUNCH();
}
// Horner combine of up to 3 sums whose window totals are ready; one block per sum
void msm_combine_stage(cudaStream_t s, const MsmPlan* const* plans, MsmWorkspace* const* wss, G1Jac* const* outs, int njobs) {
    CombineJobs cj;
    for (int j = 0; j < njobs; ++j) {
        cj.winsums[j] = wss[j]->winsums; cj.out[j] = outs[j]; cj.W[j] = plans[j]->W; cj.c[j] = plans[j]->c;
    }
    k_msm_combine<<<njobs, 32, 0, s>>>(cj);
    KZ_COUNT_LAUNCH();
}
void msm_reduce_stage_multi(cudaStream_t s, const MsmPlan* const* plans, MsmWorkspace* const* wss, G1Jac* const* outs, int njobs) {
    for (int j = 0; j < njobs; ++j) msm_window_sums_stage(s, *plans[j], *wss[j], false);
    msm_combine_stage(s, plans, wss, outs, njobs);
}
void msm_reduce_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws, G1Jac* out) {
    const MsmPlan* p = &plan;
    MsmWorkspace* w = &ws;
    msm_reduce_stage_multi(s, &p, &w, &out, 1);
}
