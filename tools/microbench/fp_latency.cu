// Latency microbenchmark (sm_100a) for the serial tail of a batch: cycles per DEPENDENT operation as seen by one
// warp on an otherwise idle SM.  clock64() around `iters` dependent repetitions, one block, result = clk / op.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../kzg_batch_verification_scheme_b200/csrc -o fp_latency fp_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "g1.cuh"

enum { M_MUL, M_SQR, M_MUL2, M_ADD, M_SUB, M_SHFL, M_SHFL12, M_SMEM, M_BAR, M_XDBL, M_XADD, M_JDBL, M_JADD, M_WIDE8, M_SWARP, M_N };
static const char* NAMES[] = {"fp_mul chain", "fp_sqr chain", "2 independent fp_mul chains / thread (per pair)", "fp_add chain", "fp_sub chain",
                              "shfl chain (1 word)", "12 independent shfl + 1 dependent", "smem Fp store -> syncwarp -> load neighbour", "__syncthreads",
                              "xyzz_dbl (1 thread)", "xyzz_add (1 thread)", "jac_dbl (1 thread)", "jac_add (1 thread)",
                              "8 independent IMAD.WIDE.X chains (per 64 wide mads)", "__syncwarp"};

__global__ void k_lat(int mode, int active_lanes, int active_warps, int warp_stride, int iters, const Fp* in, Fp* out, long long* clk) {
    __shared__ Fp sm[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool on = lane < active_lanes && (warp % warp_stride) == 0 && (warp / warp_stride) < active_warps;
    Fp a = in[0], b = in[1], c = in[2], d = in[3];
    a.v[0] ^= threadIdx.x;
    G1Xyzz P = {a, b, c, d}, Q = {b, c, d, a};
    G1Jac J = {a, b, c}, K = {c, a, b};
    uint32_t w = threadIdx.x;
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    if (mode == M_BAR) {
        for (int i = 0; i < iters; ++i) __syncthreads();
    } else if (on) {
        switch (mode) {
            case M_MUL: for (int i = 0; i < iters; ++i) a = fp_mul(a, b); break;
            case M_SQR: for (int i = 0; i < iters; ++i) a = fp_sqr(a); break;
            case M_MUL2: for (int i = 0; i < iters; ++i) { a = fp_mul(a, b); c = fp_mul(c, d); } break;
            case M_ADD: for (int i = 0; i < iters; ++i) a = fp_add(a, b); break;
            case M_SUB: for (int i = 0; i < iters; ++i) a = fp_sub(a, b); break;
            case M_SHFL: for (int i = 0; i < iters; ++i) w = __shfl_sync(0xFFFFFFFFu, w, (lane + 1) & 31) + 1; break;
            case M_SHFL12:
                for (int i = 0; i < iters; ++i) {
#pragma unroll
                    for (int k = 0; k < 12; ++k) a.v[k] = __shfl_sync(0xFFFFFFFFu, a.v[k], (lane + 1) & 31);
                    a.v[0] += a.v[11];
                }
                break;
            case M_SMEM:
                for (int i = 0; i < iters; ++i) {
                    sm[threadIdx.x] = a;
                    __syncwarp();
                    a = sm[threadIdx.x ^ 1];
                    a.v[0] += 1;
                    __syncwarp();
                }
                break;
            case M_SWARP: for (int i = 0; i < iters; ++i) { __syncwarp(); w += 1; } break;
            case M_XDBL: for (int i = 0; i < iters; ++i) P = xyzz_dbl(P); break;
            case M_XADD: for (int i = 0; i < iters; ++i) P = xyzz_add(P, Q); break;
            case M_JDBL: for (int i = 0; i < iters; ++i) J = jac_dbl(J); break;
            case M_JADD: for (int i = 0; i < iters; ++i) J = jac_add(J, K); break;
            case M_WIDE8:
                for (int i = 0; i < iters; ++i) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        uint32_t* q = &r[8 * cc];
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
                            : "r"(r[(8 * cc + 9) & 31]), "r"(w | 1u));
                    }
                }
                break;
        }
    }
    long long t1 = clock64();
    if (on) {
        uint32_t acc = w;
        for (int i = 0; i < 32; ++i) acc ^= r[i];
        a.v[1] ^= acc ^ c.v[0] ^ P.X.v[0] ^ P.ZZZ.v[3] ^ J.X.v[0] ^ J.Z.v[5];
        out[threadIdx.x] = a;
    }
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
    Fp h[4];
    for (int k = 0; k < 4; ++k) for (int i = 0; i < 12; ++i) h[k].v[i] = 0x01234567u * (i + 3 * k + 1) ^ (0x9E3779B9u >> k);
    for (int k = 0; k < 4; ++k) h[k].v[11] &= 0x0FFFFFFFu;           // < p
    Fp *din, *dout; long long* dclk;
    cudaMalloc(&din, sizeof h); cudaMalloc(&dout, sizeof(Fp) * 1024); cudaMalloc(&dclk, 8);
    cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
    struct Cfg { int lanes, warps, stride, threads; const char* what; };
    const Cfg cfgs[] = {{1, 1, 1, 32, "1 lane, 1 warp"}, {4, 1, 1, 32, "4 lanes, 1 warp"}, {32, 1, 1, 32, "32 lanes, 1 warp"},
                        {32, 4, 1, 128, "4 warps (one per SMSP)"}, {32, 2, 4, 256, "2 warps on the SAME SMSP"}, {32, 4, 4, 512, "4 warps on the SAME SMSP"},
                        {32, 8, 1, 256, "8 warps (2 per SMSP)"}};
    for (int mode = 0; mode < M_N; ++mode) {
        printf("%s\n", NAMES[mode]);
        for (const Cfg& c : cfgs) {
            if ((mode == M_BAR) && c.stride != 1) continue;
            if ((mode >= M_SHFL && mode <= M_SMEM || mode == M_SWARP) && c.lanes != 32) continue;
            int iters = (mode == M_XDBL || mode == M_XADD || mode == M_JDBL || mode == M_JADD) ? 50 : 400;
            long long best = 1LL << 62;
            for (int rep = 0; rep < 3; ++rep) {
                k_lat<<<1, c.threads>>>(mode, c.lanes, c.warps, c.stride, iters, din, dout, dclk);
                long long v = 0;
                if (cudaMemcpy(&v, dclk, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                if (v < best) best = v;
            }
            printf("    %-28s %9.1f clk/op\n", c.what, (double)best / iters);
        }
    }
    return 0;
}
