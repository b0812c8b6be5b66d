// Do the integer (IMAD.WIDE) and FP64 (DFMA) Fp multipliers overlap on one SM?  Dependent squaring chains, like
// K1a's square root: kernel A uses field.cuh (12 x u32, inline-PTX carry chains), kernel B uses fpd.cuh (8 x 48-bit
// limbs in doubles).  Measured alone and co-resident (two streams, grids sized to fit together).
// Also checks on the device that both give the same bits.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fpd.cuh"

__device__ unsigned long long g_probe[4];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); return v; }
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_int_chain(const Fp* __restrict__ in, Fp* __restrict__ out, int iters, int use_mul) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long c0 = clock64(); unsigned long long n0 = gtimer();
    Fp x = in[t], y = in[t ^ 1];
    if (use_mul) for (int i = 0; i < iters; ++i) x = fp_mul(x, y);
    else for (int i = 0; i < iters; ++i) x = fp_sqr(x);
    out[t] = x;
    if (t == 0) { g_probe[0] = clock64() - c0; g_probe[1] = gtimer() - n0; }
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_f64_chain(const Fp* __restrict__ in, Fp* __restrict__ out, int iters, int use_mul) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long c0 = clock64(); unsigned long long n0 = gtimer();
    FpD x = fpd_from_fp(in[t]), y = fpd_from_fp(in[t ^ 1]);
    if (use_mul) for (int i = 0; i < iters; ++i) x = fpd_mul(x, y);
    else for (int i = 0; i < iters; ++i) x = fpd_sqr(x);
    out[t] = fpd_to_fp(x);
    if (t == 0) { g_probe[2] = clock64() - c0; g_probe[3] = gtimer() - n0; }
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int maxthreads = sms * 8 * 128;
    Fp* h = (Fp*)malloc(sizeof(Fp) * maxthreads);
    srand(7);
    for (int i = 0; i < maxthreads; ++i) {
        for (int k = 0; k < 12; ++k) h[i].v[k] = (u32)rand() * 2654435761u + (u32)rand();
        h[i].v[11] &= 0x0FFFFFFFu;                     // below p
    }
    Fp *din, *oa, *ob;
    cudaMalloc(&din, sizeof(Fp) * maxthreads); cudaMalloc(&oa, sizeof(Fp) * maxthreads); cudaMalloc(&ob, sizeof(Fp) * maxthreads);
    cudaMemcpy(din, h, sizeof(Fp) * maxthreads, cudaMemcpyHostToDevice);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    // ---- correctness on the device: same bits from both multipliers
    for (int um = 0; um < 2; ++um) {
        k_int_chain<3><<<sms, 128>>>(din, oa, 100, um);
        k_f64_chain<2><<<sms, 128>>>(din, ob, 100, um);
        cudaDeviceSynchronize();
        Fp* ra = (Fp*)malloc(sizeof(Fp) * sms * 128); Fp* rb = (Fp*)malloc(sizeof(Fp) * sms * 128);
        cudaMemcpy(ra, oa, sizeof(Fp) * sms * 128, cudaMemcpyDeviceToHost); cudaMemcpy(rb, ob, sizeof(Fp) * sms * 128, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < sms * 128; ++i) for (int k = 0; k < 12; ++k) if (ra[i].v[k] != rb[i].v[k]) { ++bad; break; }
        printf("check %s chain x100: %d mismatches of %d  (%s)\n", um ? "mul" : "sqr", bad, sms * 128, cudaGetErrorString(cudaGetLastError()));
        free(ra); free(rb);
    }
    const int iters = 4000;
    double last_mhz[2] = {0, 0};
    auto run = [&](int bpsA, int bpsB, int itA, int itB, int um) {
        // bps = blocks per SM (0 = kernel not launched)
        cudaDeviceSynchronize();
        double t0 = now_ms();
        if (bpsA) { if (bpsA >= 3) k_int_chain<3><<<sms * bpsA, 128, 0, s1>>>(din, oa, itA, um); else k_int_chain<2><<<sms * bpsA, 128, 0, s1>>>(din, oa, itA, um); }
        if (bpsB) { if (bpsB >= 3) k_f64_chain<3><<<sms * bpsB, 128, 0, s2>>>(din, ob, itB, um); else if (bpsB == 2) k_f64_chain<2><<<sms * bpsB, 128, 0, s2>>>(din, ob, itB, um); else k_f64_chain<1><<<sms * bpsB, 128, 0, s2>>>(din, ob, itB, um); }
        cudaDeviceSynchronize();
        double dt = now_ms() - t0;
        unsigned long long pr[4]; cudaMemcpyFromSymbol(pr, g_probe, sizeof(pr));
        last_mhz[0] = pr[1] ? (double)pr[0] / (double)pr[1] * 1e3 : 0; last_mhz[1] = pr[3] ? (double)pr[2] / (double)pr[3] * 1e3 : 0;
        return dt;
    };
    for (int um = 0; um < 2; ++um) {
        const char* nm = um ? "mul" : "sqr";
        run(3, 0, 100, 0, um); run(0, 2, 0, 100, um);
        double best[8] = {0};
        for (int bA = 1; bA <= 4; ++bA) {
            double ms = run(bA, 0, iters, 0, um);
            double rate = (double)sms * bA * 128 * iters / (ms * 1e-3);
            best[bA] = rate;
            printf("%s INT  alone %d blk/SM: %8.3f ms  %7.2f G%s/s  %.2f clk/SM per %s [MHz %.0f]\n", nm, bA, ms, rate / 1e9, nm, sms * (clk * 1e3) / rate, nm, last_mhz[0]);
        }
        double bestB[8] = {0};
        for (int bB = 1; bB <= 4; ++bB) {
            double ms = run(0, bB, 0, iters, um);
            double rate = (double)sms * bB * 128 * iters / (ms * 1e-3);
            bestB[bB] = rate;
            printf("%s F64  alone %d blk/SM: %8.3f ms  %7.2f G%s/s  %.2f clk/SM per %s [MHz %.0f]\n", nm, bB, ms, rate / 1e9, nm, sms * (clk * 1e3) / rate, nm, last_mhz[1]);
        }
        // co-resident: choose iteration counts so that both would take about the same time at their stand-alone rates
        int combos[6][2] = {{2, 1}, {2, 2}, {3, 1}, {1, 2}, {3, 2}, {1, 1}};
        for (auto& c : combos) {
            int bA = c[0], bB = c[1];
            double perA = best[bA] / (sms * bA * 128), perB = bestB[bB] / (sms * bB * 128);   // iterations/s per thread
            for (int pass = 0; pass < 2; ++pass) {
                int itA = iters, itB = (int)(iters * perB / perA);
                if (pass == 1) { itB = (int)(itB * 0.6); }                 // second guess: contention slows B more than A
                double ms = run(bA, bB, itA, itB, um);
                double total = (double)sms * 128 * ((double)bA * itA + (double)bB * itB);
                printf("%s MIX INT %d + F64 %d blk/SM (it %d/%d): %8.3f ms  %7.2f G%s/s  = %.2fx best INT alone  [SM MHz seen by INT %.0f, F64 %.0f]\n", nm, bA, bB, itA, itB, ms,
                       total / (ms * 1e-3) / 1e9, nm, total / (ms * 1e-3) / fmax(fmax(best[1], best[2]), fmax(best[3], best[4])), last_mhz[0], last_mhz[1]);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
