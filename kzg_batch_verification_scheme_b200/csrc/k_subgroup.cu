// K1b / K1c: the two |x| double-and-add chains of the subgroup check (see k_decompress.cu for the split).
// (An out-of-line build of the Fp products -- one 6 KB copy per kernel so that the doubling loop fits the
// instruction cache -- was measured: 101 ms vs 98 ms inlined at n = 2^20, so the products stay inlined.)
#include "kernels.h"

__device__ __forceinline__ Fp ld_fp(const Fp* p) {
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4 a = s[0], b = s[1], c = s[2];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
    return r;
}
__device__ __forceinline__ void st_fp(Fp* p, const Fp& r) {
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    d[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    d[2] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_subgroup_chain1(const Fp* __restrict__ pts, size_t m, Fp* __restrict__ tmp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    G1Aff p = {ld_fp(pts + 2 * i), ld_fp(pts + 2 * i + 1)};
    if (aff_is_inf(p)) return;                       // infinity or already rejected: nothing to check
    G1Jac t = jac_mul_xabs_aff(p);
    st_fp(tmp + 3 * i, t.X); st_fp(tmp + 3 * i + 1, t.Y); st_fp(tmp + 3 * i + 2, t.Z);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_subgroup_chain2(Fp* __restrict__ pts, size_t m, const Fp* __restrict__ tmp,
                                                               u8* __restrict__ status, u32* __restrict__ counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fp px = ld_fp(pts + 2 * i);
    {
        Fp py = ld_fp(pts + 2 * i + 1);
        if (fp_is_zero(px) && fp_is_zero(py)) return;
    }
    G1Jac t = {ld_fp(tmp + 3 * i), ld_fp(tmp + 3 * i + 1), ld_fp(tmp + 3 * i + 2)};
    G1Jac q = jac_mul_xabs(t);
    G1Aff p = {px, ld_fp(pts + 2 * i + 1)};
    if (!g1_subgroup_compare(p, q)) {
        Fp z = fp_zero();
        st_fp(pts + 2 * i, z); st_fp(pts + 2 * i + 1, z);
        status[i] = (u8)ST_NOT_IN_G1;
        atomicAdd(counters, 1u);
    }
}


void launch_subgroup_chains(cudaStream_t s, Fp* pts, size_t m, Fp* tmp, uint8_t* status, uint32_t* counters, int mb, int mc) {
    unsigned blocks = (unsigned)((m + 127) / 128);
    if (mb >= 4) k_subgroup_chain1<4><<<blocks, 128, 0, s>>>(pts, m, tmp);
    else if (mb == 3) k_subgroup_chain1<3><<<blocks, 128, 0, s>>>(pts, m, tmp);
    else k_subgroup_chain1<2><<<blocks, 128, 0, s>>>(pts, m, tmp);
    if (mc >= 4) k_subgroup_chain2<4><<<blocks, 128, 0, s>>>(pts, m, tmp, status, counters);
    else if (mc == 3) k_subgroup_chain2<3><<<blocks, 128, 0, s>>>(pts, m, tmp, status, counters);
    else k_subgroup_chain2<2><<<blocks, 128, 0, s>>>(pts, m, tmp, status, counters);
}

// ---- batched subgroup check (msm.cuh "batched subgroup check"): 128 slice sums of one MSM's buckets.
// A window with k magnitude bits is a 2^kh x 2^kl table of buckets (m = row * 2^kl + col, kl = k / 2).  Bit b of
// m is a column bit (b < kl) or a row bit, so every bit slice is a sum of column totals or of row totals:
//   pass 1  partial sums of runs of KZ_SG_RUN elements along rows and along columns (every lane busy)
//   pass 2  row totals R[r] and column totals Q[c]
//   pass 3  slice b = sum of Q[c] with bit b of c, or of R[r] with bit b - kl of r; "all" slice = sum of R
//           plus the bucket of magnitude 2^k; then one thread runs the |x|^2 chain on the slice sum.
// 2 * 2^k additions per window instead of (k + 1) * 2^(k-1) for slice-by-slice reductions.
#define KZ_SG_RUN 16
struct SgWin { int k, kl, kh; u32 rows, cols, tpr, tpc, rjobs, jobs; };
__device__ __forceinline__ SgWin sg_win(const MsmPlan& P, int w) {
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
__device__ __forceinline__ G1Xyzz sg_bucket(const G1Xyzz* base, u32 m) { return m ? base[m - 1] : xyzz_inf(); }
// part: [W][stride] partial sums (row jobs first, then column jobs)
__global__ void __launch_bounds__(128) k_sg_pass1(const G1Xyzz* __restrict__ buckets, G1Xyzz* __restrict__ part, u32 stride,
                                                   const __grid_constant__ MsmPlan plan) {
    const int w = blockIdx.y;
    const SgWin g = sg_win(plan, w);
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.jobs) return;
    const G1Xyzz* base = buckets + plan.bucket_off[w];
    G1Xyzz acc = xyzz_inf();
    if (t < g.rjobs) {
        u32 r = t / g.tpr, p = t - r * g.tpr, len = g.cols / g.tpr;
        u32 m0 = (r << g.kl) + p * len;
        for (u32 i = 0; i < len; ++i) acc = xyzz_add(acc, sg_bucket(base, m0 + i));
    } else {
        u32 u = t - g.rjobs, c = u / g.tpc, p = u - c * g.tpc, len = g.rows / g.tpc;
        for (u32 i = 0; i < len; ++i) acc = xyzz_add(acc, sg_bucket(base, ((p * len + i) << g.kl) + c));
    }
    part[(size_t)w * stride + t] = acc;
}
// tot: [W][tstride]: R[0..rows) then Q[0..cols)
__global__ void __launch_bounds__(128) k_sg_pass2(const G1Xyzz* __restrict__ part, u32 stride, G1Xyzz* __restrict__ tot, u32 tstride,
                                                   const __grid_constant__ MsmPlan plan) {
    const int w = blockIdx.y;
    const SgWin g = sg_win(plan, w);
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.rows + g.cols) return;
    const G1Xyzz* src = part + (size_t)w * stride + (t < g.rows ? t * g.tpr : g.rjobs + (t - g.rows) * g.tpc);
    const u32 cnt = t < g.rows ? g.tpr : g.tpc;
    G1Xyzz acc = src[0];
    for (u32 i = 1; i < cnt; ++i) acc = xyzz_add(acc, src[i]);
    tot[(size_t)w * tstride + t] = acc;
}
__global__ void __launch_bounds__(32) k_sg_pass3(const G1Xyzz* __restrict__ buckets, const G1Xyzz* __restrict__ tot, u32 tstride,
                                                  u32* __restrict__ counters, const __grid_constant__ MsmPlan plan) {
    __shared__ G1Xyzz red[32];
    const SgSlice sl = sg_slice(plan, (int)blockIdx.x);
    const SgWin g = sg_win(plan, sl.w);
    const G1Xyzz* R = tot + (size_t)sl.w * tstride;
    const G1Xyzz* Q = R + g.rows;
    G1Xyzz acc = xyzz_inf();
    if (sl.all) {
        for (u32 r = threadIdx.x; r < g.rows; r += 32) acc = xyzz_add(acc, R[r]);
        if (threadIdx.x == 0) acc = xyzz_add(acc, buckets[plan.bucket_off[sl.w] + (1u << g.k) - 1u]);   // magnitude 2^k
    } else if (sl.b < g.kl) {
        for (u32 c = threadIdx.x; c < g.cols; c += 32) if ((c >> sl.b) & 1u) acc = xyzz_add(acc, Q[c]);
    } else {
        for (u32 r = threadIdx.x; r < g.rows; r += 32) if ((r >> (sl.b - g.kl)) & 1u) acc = xyzz_add(acc, R[r]);
    }
    red[threadIdx.x] = acc;
    __syncwarp();
    for (int st = 16; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) red[threadIdx.x] = xyzz_add(red[threadIdx.x], red[threadIdx.x + st]);
        __syncwarp();
    }
    if (threadIdx.x == 0 && !sg_sum_in_g1(red[0])) atomicAdd(counters + 2, 1u);
}
// counters[2] += number of slice sums outside G1.  work: KZ_SG_WORK_ENTRIES G1Xyzz of scratch.
void launch_sg_batch_check(cudaStream_t s, const MsmPlan& plan, const G1Xyzz* buckets, G1Xyzz* work, uint32_t* counters) {
    int kmax = plan.c - 1, tb = plan.nbits - plan.c * (plan.W - 1);
    if (plan.W == 1 || tb > kmax) kmax = tb;
    int kl = kmax / 2, kh = kmax - kl;
    // kl = floor(k/2) and kh = ceil(k/2) are monotone in k, so the widest window bounds every stride
    u32 rows = 1u << kh, cols = 1u << kl;
    u32 tpr = cols > KZ_SG_RUN ? cols / KZ_SG_RUN : 1u, tpc = rows > KZ_SG_RUN ? rows / KZ_SG_RUN : 1u;
    u32 stride = rows * tpr + cols * tpc, tstride = rows + cols;
    G1Xyzz* part = work;
    G1Xyzz* tot = work + (size_t)plan.W * stride;
    k_sg_pass1<<<dim3((stride + 127) / 128, (unsigned)plan.W), 128, 0, s>>>(buckets, part, stride, plan);
    KZ_COUNT_LAUNCH();
    k_sg_pass2<<<dim3((tstride + 127) / 128, (unsigned)plan.W), 128, 0, s>>>(part, stride, tot, tstride, plan);
    KZ_COUNT_LAUNCH();
    k_sg_pass3<<<(unsigned)plan.nbits, 32, 0, s>>>(buckets, tot, tstride, counters, plan);
    KZ_COUNT_LAUNCH();
}
size_t sg_work_entries(const MsmPlan& plan) {
    int kmax = plan.c - 1, tb = plan.nbits - plan.c * (plan.W - 1);
    if (plan.W == 1 || tb > kmax) kmax = tb;
    int kl = kmax / 2, kh = kmax - kl;
    u32 rows = 1u << kh, cols = 1u << kl;
    u32 tpr = cols > KZ_SG_RUN ? cols / KZ_SG_RUN : 1u, tpc = rows > KZ_SG_RUN ? rows / KZ_SG_RUN : 1u;
    return (size_t)plan.W * (rows * tpr + cols * tpc + rows + cols);
}
