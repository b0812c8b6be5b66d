// K1: batch G1 decompression + on-curve + subgroup check (BASELINE.json:5 item (b)) -- ~89 % of all
// integer work of a batch verification.  One thread per point; 48-byte inputs are
// read with three 128-bit loads, the 96-byte Montgomery affine result is written with six.
// Also hosts the small conversion kernels, the primitive debug operator and the IMAD microbenchmark.
#include "kernels.h"

std::atomic<uint64_t> g_kzgb_launches{0};

__device__ __forceinline__ u32 ld_be32(u32 x) { return __byte_perm(x, 0, 0x0123); }

__global__ void __launch_bounds__(128) k_decompress(const u8* __restrict__ inC, const u8* __restrict__ inPi, size_t n,
                                                    Fp* __restrict__ out_pts, u8* __restrict__ status,
                                                    u32* __restrict__ counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    const u8* src = i < n ? inC + 48 * i : inPi + 48 * (i - n);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4 q0 = __ldg(s4), q1 = __ldg(s4 + 1), q2 = __ldg(s4 + 2);
    u32 w[12] = {ld_be32(q0.x), ld_be32(q0.y), ld_be32(q0.z), ld_be32(q0.w), ld_be32(q1.x), ld_be32(q1.y),
                 ld_be32(q1.z), ld_be32(q1.w), ld_be32(q2.x), ld_be32(q2.y), ld_be32(q2.z), ld_be32(q2.w)};
    G1Aff p;
    u32 st = g1_decompress_validate(p, w);
    uint4* d4 = reinterpret_cast<uint4*>(out_pts + 2 * i);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d4[k] = make_uint4(p.x.v[4 * k], p.x.v[4 * k + 1], p.x.v[4 * k + 2], p.x.v[4 * k + 3]);
        d4[3 + k] = make_uint4(p.y.v[4 * k], p.y.v[4 * k + 1], p.y.v[4 * k + 2], p.y.v[4 * k + 3]);
    }
    status[i] = (u8)st;
    if (st) atomicAdd(counters, 1u);
}

void launch_decompress(cudaStream_t s, const uint8_t* inC, const uint8_t* inPi, size_t n, Fp* out_pts, uint8_t* status,
                       uint32_t* counters) {
    if (!n) return;
    size_t blocks = (2 * n + 127) / 128;
    k_decompress<<<(unsigned)blocks, 128, 0, s>>>(inC, inPi, n, out_pts, status, counters);
    KZ_COUNT_LAUNCH();
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

// ---- IMAD.WIDE.U32 issue-rate microbenchmark: 8 independent carry-less chains per thread.
// Each loop iteration issues 8 x 16 wide multiply-adds; total per thread = iters * 128.
__global__ void k_imad_bench(u32* sink, int iters) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u64 a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
    u32 x = t | 1, y = (t * 2654435761u) | 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a4) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a5) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a6) : "r"(x), "r"(y));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a7) : "r"(x), "r"(y));
        }
    }
    u64 r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x123456789ull) sink[0] = (u32)r;      // never true in practice; keeps the chains alive
}
void launch_imad_bench(cudaStream_t s, uint32_t* sink, int blocks, int threads, int iters) {
    k_imad_bench<<<blocks, threads, 0, s>>>(sink, iters);
    KZ_COUNT_LAUNCH();
}
