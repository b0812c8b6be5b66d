// K1: batch G1 decompression + on-curve + subgroup check (BASELINE.json:5 item (b)).  One thread per point; 48-byte
// inputs are read with three 128-bit loads, the 96-byte Montgomery affine result is written with six.
// Batches of two or more proofs run K1a alone and establish subgroup membership on the bucket-slice sums of the
// MSMs (msm.cuh, DESIGN.md "Batched subgroup check"); K1b/K1c are the per-point check used by verify_kzg_proof,
// the cell commitments, kzgb_g1_decompress_batch and as the fallback that names the offending points.
// Also hosts the small conversion kernels, the primitive debug operator and the IMAD microbenchmark.
#include <cstdlib>

#include "kernels.h"

std::atomic<uint64_t> g_kzgb_launches{0};

__device__ __forceinline__ u32 ld_be32(u32 x) { return __byte_perm(x, 0, 0x0123); }

// K1 is split into three kernels so that each stays within a register budget that allows 12-16 resident
// warps per SM (the fused kernel needed 255 registers = 8 warps and was latency-bound, profiles/):
//   K1a  flags, range check, y = sqrt(x^3+4), sign            -> affine P            (~480 Fp products)
//   K1b  T = [|x|]P, Jacobian double-and-add, mixed additions -> T (144 B scratch)   (~500)
//   K1c  Q = [|x|]T, full additions; sigma(P) == -Q ?         -> status, P or zeros  (~525)
// The 240 B/point of extra traffic is ~0.1 % of the kernels' run time.
// MINB = minimum resident blocks per SM the register allocation must allow (run-time pick: KZGB_K1_MINB).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_decompress_sqrt(const u8* __restrict__ inC, const u8* __restrict__ inPi, size_t n,
                                                               size_t total, Fp* __restrict__ out_pts, u8* __restrict__ status,
                                                               u32* __restrict__ counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const u8* src = i < n ? inC + 48 * i : inPi + 48 * (i - n);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4 q0 = __ldg(s4), q1 = __ldg(s4 + 1), q2 = __ldg(s4 + 2);
    u32 w[12] = {ld_be32(q0.x), ld_be32(q0.y), ld_be32(q0.z), ld_be32(q0.w), ld_be32(q1.x), ld_be32(q1.y),
                 ld_be32(q1.z), ld_be32(q1.w), ld_be32(q2.x), ld_be32(q2.y), ld_be32(q2.z), ld_be32(q2.w)};
    G1Aff p;
    u32 st = g1_decompress_sqrt(p, w);
    uint4* d4 = reinterpret_cast<uint4*>(out_pts + 2 * i);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d4[k] = make_uint4(p.x.v[4 * k], p.x.v[4 * k + 1], p.x.v[4 * k + 2], p.x.v[4 * k + 3]);
        d4[3 + k] = make_uint4(p.y.v[4 * k], p.y.v[4 * k + 1], p.y.v[4 * k + 2], p.y.v[4 * k + 3]);
    }
    status[i] = (u8)st;
    if (st) atomicAdd(counters, 1u);
}
// per-kernel occupancy choice "abc" (digits 2..4 for K1a, K1b, K1c), e.g. KZGB_K1_MINB=433
// points [0, total) with total <= 2n: i < n from inC, the rest from inPi
static int k1_minb_cfg() {
    static const int cfg = [] { const char* e = getenv("KZGB_K1_MINB"); int v = e ? atoi(e) : 322; return (v >= 222 && v <= 444) ? v : 322; }();
    return cfg;
}
// with_subgroup = false: K1a only (flags, x < p, on-curve, y); the caller runs the batched subgroup check
static void launch_decompress_impl(cudaStream_t s, const uint8_t* inC, const uint8_t* inPi, size_t n, size_t total, Fp* out_pts,
                                   Fp* tmp, uint8_t* status, uint32_t* counters, bool with_subgroup = true) {
    if (!total) return;
    const int cfg = k1_minb_cfg();
    const int ma = cfg / 100, mb = cfg / 10 % 10, mc = cfg % 10;
    unsigned blocks = (unsigned)((total + 127) / 128);
    if (ma >= 4) k_decompress_sqrt<4><<<blocks, 128, 0, s>>>(inC, inPi, n, total, out_pts, status, counters);
    else if (ma == 3) k_decompress_sqrt<3><<<blocks, 128, 0, s>>>(inC, inPi, n, total, out_pts, status, counters);
    else k_decompress_sqrt<2><<<blocks, 128, 0, s>>>(inC, inPi, n, total, out_pts, status, counters);
    KZ_COUNT_LAUNCH();
    if (!with_subgroup) return;
    launch_subgroup_chains(s, out_pts, total, tmp, status, counters, mb, mc);
    KZ_COUNT_LAUNCH(); KZ_COUNT_LAUNCH();
}
// per-point subgroup check of already decompressed points (the fallback of the batched check)
void launch_subgroup_points(cudaStream_t s, Fp* pts, size_t m, Fp* tmp, uint8_t* status, uint32_t* counters) {
    if (!m) return;
    const int cfg = k1_minb_cfg();
    launch_subgroup_chains(s, pts, m, tmp, status, counters, cfg / 10 % 10, cfg % 10);
    KZ_COUNT_LAUNCH(); KZ_COUNT_LAUNCH();
}
// K1a only on both input arrays in one launch (small batches: two launches would be two serial latencies)
void launch_decompress_sqrt(cudaStream_t s, const uint8_t* inC, const uint8_t* inPi, size_t n, Fp* out_pts, uint8_t* status,
                            uint32_t* counters) {
    launch_decompress_impl(s, inC, inPi, n, 2 * n, out_pts, nullptr, status, counters, false);
}
void launch_decompress_sqrt_points(cudaStream_t s, const uint8_t* in, size_t m, Fp* out_pts, uint8_t* status, uint32_t* counters) {
    launch_decompress_impl(s, in, in, m, m, out_pts, nullptr, status, counters, false);
}
void launch_decompress(cudaStream_t s, const uint8_t* inC, const uint8_t* inPi, size_t n, Fp* out_pts, Fp* tmp, uint8_t* status,
                       uint32_t* counters) {
    launch_decompress_impl(s, inC, inPi, n, 2 * n, out_pts, tmp, status, counters);
}
// m points of ONE input array (used per half so that K1 on the commitments starts while pi, z, y still copy)
void launch_decompress_points(cudaStream_t s, const uint8_t* in, size_t m, Fp* out_pts, Fp* tmp, uint8_t* status,
                              uint32_t* counters) {
    launch_decompress_impl(s, in, in, m, m, out_pts, tmp, status, counters);
}

// ---- conversions between device Montgomery affine and canonical big-endian bytes
__global__ void k_points_to_be(const Fp* __restrict__ pts, size_t m, u8* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    G1Aff p = {pts[2 * i], pts[2 * i + 1]};
    aff_to_be96(out + 96 * i, p);
}
void launch_points_to_be(cudaStream_t s, const Fp* pts, size_t m, uint8_t* out96) {
    if (!m) return;
    k_points_to_be<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(pts, m, out96);
    KZ_COUNT_LAUNCH();
}
__global__ void k_points_from_be(const u8* __restrict__ in, size_t m, Fp* __restrict__ pts, u32* counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    G1Aff p;
    bool ok = aff_from_be96(p, in + 96 * i);
    if (ok && !aff_is_inf(p)) {      // must satisfy the curve equation
        Fp rhs = fp_add(fp_mul(fp_sqr(p.x), p.x), fp_const(FP_B));
        ok = fp_eq(fp_sqr(p.y), rhs);
    }
    if (!ok) { p = aff_inf(); atomicAdd(counters, 1u); }
    pts[2 * i] = p.x;
    pts[2 * i + 1] = p.y;
}
void launch_points_from_be(cudaStream_t s, const uint8_t* in96, size_t m, Fp* pts, uint32_t* counters) {
    if (!m) return;
    k_points_from_be<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(in96, m, pts, counters);
    KZ_COUNT_LAUNCH();
}

// ---- primitive debug operator (tests): one thread per record
__device__ __noinline__ Fp d_fp_inv(const Fp& a) { return fp_inv(a); }

__global__ void k_debug_op(int op, const u8* __restrict__ in, u8* __restrict__ out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    switch (op) {
        case 1: case 3: case 4: {
            Fp a, b;
            fp_from_be(a, in + 96 * i); fp_from_be(b, in + 96 * i + 48);
            Fp r = op == 1 ? fp_mul(a, b) : (op == 3 ? fp_add(a, b) : fp_sub(a, b));
            fp_to_be(out + 48 * i, r);
            break;
        }
        case 2: case 5: case 6: {
            Fp a;
            fp_from_be(a, in + 48 * i);
            Fp r = op == 2 ? fp_sqr(a) : (op == 5 ? d_fp_inv(a) : fp_sqrt_candidate(a));
            fp_to_be(out + 48 * i, r);
            break;
        }
        case 7: case 8: {
            Fr a, b;
            fr_raw_from_be(a, in + 64 * i); fr_raw_from_be(b, in + 64 * i + 32);
            a = fr_to_mont(a); b = fr_to_mont(b);
            Fr r = op == 7 ? fr_mul(a, b) : fr_add(a, b);
            fr_raw_to_be(out + 32 * i, fr_from_mont(r));
            break;
        }
        case 9: {
            G1Aff p, q;
            aff_from_be96(p, in + 192 * i); aff_from_be96(q, in + 192 * i + 96);
            G1Jac r = jac_add(jac_from_aff(p), jac_from_aff(q));
            aff_to_be96(out + 96 * i, jac_to_aff(r));
            break;
        }
        case 10: case 12: {
            G1Aff p;
            aff_from_be96(p, in + 96 * i);
            G1Jac r;
            if (aff_is_inf(p)) r = jac_inf();
            else r = op == 10 ? jac_dbl(jac_from_aff(p)) : jac_mul_xabs(jac_mul_xabs_aff(p));
            aff_to_be96(out + 96 * i, jac_to_aff(r));
            break;
        }
        case 11: {
            G1Aff p;
            aff_from_be96(p, in + 128 * i);
            Fr k;
            fr_raw_from_be(k, in + 128 * i + 96);
            G1Jac r = jac_mul_limbs(jac_from_aff(p), k.v, 8);
            aff_to_be96(out + 96 * i, jac_to_aff(r));
            break;
        }
        default: break;
    }
}
void launch_debug_op(cudaStream_t s, int op, const uint8_t* in, uint8_t* out, size_t count) {
    if (!count) return;
    k_debug_op<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(op, in, out, count);
    KZ_COUNT_LAUNCH();
}

// ---- integer multiply-add issue-rate microbenchmarks (roofline denominators, measured on the box).
// mode 0: carry-chained mad.lo.cc/madc.hi.cc pairs -> IMAD.WIDE.U32.X, the instruction every Montgomery
//         product in this library is made of (8 independent 16-long chains per loop trip and thread);
// mode 1: plain 32-bit mad.lo.u32 -> IMAD (the pipe's nominal issue rate).
// Multiplicands are loop-carried registers so ptxas cannot hoist the products out of the loop.
// Each loop trip issues 128 multiply-adds per thread in both modes.
__global__ void k_imad_wide_bench(u32* sink, int iters) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 y = (t * 2654435761u) | 1;
    u32 r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = t + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                u32* q = &r[8 * c];
                asm volatile(
                    "mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
                    "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
                    "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
                    "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
                    "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
                    "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.cc.u32 %7,%8,%9,%7;"
                    "madc.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
                    "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                    : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7])
                    : "r"(r[(8 * c + 9) & 31]), "r"(y)
                    : "memory");                     // keep each 16-long chain contiguous in the schedule
            }
        }
    }
    u32 z = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) z ^= r[i];
    if (z == 0x23456789u) sink[0] = z;      // practically never; keeps the chains alive
}
__global__ void k_imad32_bench(u32* sink, int iters) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 y = (t * 2654435761u) | 1;
    u32 r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = t + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 32; i++) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(r[i]) : "r"(r[(i + 1) & 31]), "r"(y));
        }
    }
    u32 z = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) z ^= r[i];
    if (z == 0x23456789u) sink[0] = z;
}
void launch_imad_bench(cudaStream_t s, uint32_t* sink, int blocks, int threads, int iters, int mode) {
    if (mode == 0) k_imad_wide_bench<<<blocks, threads, 0, s>>>(sink, iters);
    else k_imad32_bench<<<blocks, threads, 0, s>>>(sink, iters);
    KZ_COUNT_LAUNCH();
}
