// Phase timing of one Fp12 product round (mpair.cuh) on an idle SM: clock64 stamps of thread 0 (and of the last
// active product lane) around  products | barrier | fold | barrier, averaged over `iters` dependent rounds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../kzg_batch_verification_scheme_b200/csrc -o mp_round mp_round.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "mpair.cuh"

__global__ void __launch_bounds__(128) k_round(const Fp12* in, Fp12* out, int iters, long long* clk) {
    __shared__ MpScratch S;
    const int t = threadIdx.x, k = t >> 1;
    mp_unit_init(S.U);
    if (t < 12) { if (t & 1) S.f.c[k].c1 = in[0].c[k].c1; else S.f.c[k].c0 = in[0].c[k].c0; }
    if (t < 12) { if (t & 1) S.fb.c[k].c1 = in[1].c[k].c1; else S.fb.c[k].c0 = in[1].c[k].c0; }
    __syncthreads();
    long long acc[4] = {0, 0, 0, 0};
    for (int i = 0; i < iters; ++i) {
        long long t0 = clock64();
        mp_mul_products(S.U, t, S.f, S.fb);
        long long t1 = clock64();
        __syncthreads();
        long long t2 = clock64();
        mp_mul_fold(S.U, t, S.f);
        long long t3 = clock64();
        __syncthreads();
        long long t4 = clock64();
        acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3;
    }
    long long w0 = clock64();
    for (int i = 0; i < iters; ++i) mp_mul(S.U, S.f, S.f, S.fb);       // the out-of-line form the serial sequences call
    long long w1 = clock64();
    if (t == 0 || t == 107) { for (int j = 0; j < 4; ++j) clk[(t ? 5 : 0) + j] = acc[j]; clk[(t ? 5 : 0) + 4] = w1 - w0; }
    if (t < 12) { if (t & 1) out[0].c[k].c1 = S.f.c[k].c1; else out[0].c[k].c0 = S.f.c[k].c0; }
}

int main() {
    Fp12 h[2];
    unsigned* w = (unsigned*)h;
    for (size_t i = 0; i < sizeof h / 4; ++i) w[i] = 0x9E3779B9u * (unsigned)(i + 1);
    for (int e = 0; e < 2; ++e) for (int c = 0; c < 6; ++c) { h[e].c[c].c0.v[11] &= 0x0FFFFFFFu; h[e].c[c].c1.v[11] &= 0x0FFFFFFFu; }
    Fp12 *din, *dout; long long* dclk;
    cudaMalloc(&din, sizeof h); cudaMalloc(&dout, sizeof(Fp12)); cudaMalloc(&dclk, 80);
    cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 200;
    for (int rep = 0; rep < 2; ++rep) k_round<<<1, 128>>>(din, dout, iters, dclk);
    long long c[10];
    if (cudaMemcpy(c, dclk, 80, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
    const char* nm[5] = {"products", "barrier 1", "fold", "barrier 2", "mp_mul (out of line, whole round)"};
    for (int who = 0; who < 2; ++who) {
        printf("thread %d\n", who ? 107 : 0);
        for (int j = 0; j < 5; ++j) printf("    %-36s %8.1f clk\n", nm[j], (double)c[5 * who + j] / iters);
    }
    return 0;
}
