// Cell batch kernels (BASELINE.json config[4]): twiddle table, cell leaf hashes, per-opening iNTT-64 + scaling,
// column sums of the interpolation coefficients, per-commitment challenge weights.
#include "cells.cuh"
#include "kernels.h"

__global__ void k_cell_twiddles(Fr* __restrict__ W) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= KZ_N_EXT) return;
    Fr base = fr_const(FR_OMEGA_INV), acc = fr_const(FR_ONE);
    for (int i = 12; i >= 0; --i) {
        acc = fr_mul(acc, acc);
        if ((t >> i) & 1u) acc = fr_mul(acc, base);
    }
    W[t] = acc;
}
void launch_cell_twiddles(cudaStream_t s, Fr* W) {
    k_cell_twiddles<<<KZ_N_EXT / 128, 128, 0, s>>>(W);
    KZ_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(64) k_cell_leaf_hash(const u32* __restrict__ ci, const u32* __restrict__ xi,
                                                       const u8* __restrict__ cells, const u8* __restrict__ proofs, size_t m,
                                                       u32* __restrict__ leaves) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    u32 h[8];
    fs_cell_leaf_words(h, ci[k], xi[k], reinterpret_cast<const u32*>(cells + 2048 * k), reinterpret_cast<const u32*>(proofs + 48 * k));
    for (int i = 0; i < 8; ++i) leaves[8 * k + i] = h[i];
}
void launch_cell_leaf_hash(cudaStream_t s, const uint32_t* ci, const uint32_t* xi, const uint8_t* cells, const uint8_t* proofs, size_t m,
                           uint32_t* leaves) {
    if (!m) return;
    k_cell_leaf_hash<<<(unsigned)((m + 63) / 64), 64, 0, s>>>(ci, xi, cells, proofs, m, leaves);
    KZ_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(KZ_CELL_LEN) k_cell_scalars(const Fr* __restrict__ W, const u32* __restrict__ root_words,
                                                              const u32* __restrict__ ci, const u32* __restrict__ xi, u32 nc,
                                                              const u8* __restrict__ cells, size_t m, u64 k0, Fr* __restrict__ coefs,
                                                              u32* __restrict__ r_out, u32* __restrict__ rh_out,
                                                              u32* __restrict__ counters) {
    __shared__ CellScratch S;
    size_t k = blockIdx.x;
    if (k >= m) return;
    // k0: index of this shard's first opening in the whole batch (the challenge r_k hashes the global index)
    coop_cell_body(S, W, root_words, k0 + (u64)k, ci[k], xi[k], nc, cells + 2048 * k, coefs + KZ_CELL_LEN * k, r_out + 4 * k, rh_out + 8 * k);
    if (threadIdx.x == 0 && S.bad) atomicAdd(counters + 1, S.bad);
}
void launch_cell_scalars(cudaStream_t s, const Fr* W, const uint32_t* root_words, const uint32_t* ci, const uint32_t* xi, uint32_t nc,
                         const uint8_t* cells, size_t m, uint64_t k0, Fr* coefs, uint32_t* r_out, uint32_t* rh_out, uint32_t* counters) {
    if (!m) return;
    k_cell_scalars<<<(unsigned)m, KZ_CELL_LEN, 0, s>>>(W, root_words, ci, xi, nc, cells, m, k0, coefs, r_out, rh_out, counters);
    KZ_COUNT_LAUNCH();
}

// S_i = sum_k coefs[k][i]; writes -S_i as a canonical 8-limb scalar to out + 8 i.  One block per coefficient.
#define KZ_COL_THREADS 256
__global__ void __launch_bounds__(KZ_COL_THREADS) k_cell_column_sum(const Fr* __restrict__ coefs, size_t m, u32* __restrict__ out) {
    __shared__ Fr red[KZ_COL_THREADS];
    int i = blockIdx.x;
    Fr acc = fr_zero();
    for (size_t k = threadIdx.x; k < m; k += KZ_COL_THREADS) acc = fr_add(acc, coefs[KZ_CELL_LEN * k + i]);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = KZ_COL_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fr_add(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Fr neg = fr_from_mont(fr_neg(red[0]));
        for (int j = 0; j < 8; ++j) out[8 * i + j] = neg.v[j];
    }
}
// w_i = sum_{k: ci[k] = i} r_k  (canonical limbs in, canonical 8 limbs out).  One block per commitment.
__global__ void __launch_bounds__(KZ_COL_THREADS) k_cell_commit_weights(const u32* __restrict__ ci, const u32* __restrict__ r, size_t m,
                                                                        u32* __restrict__ out) {
    __shared__ Fr red[KZ_COL_THREADS];
    u32 i = blockIdx.x;
    Fr acc = fr_zero();
    for (size_t k = threadIdx.x; k < m; k += KZ_COL_THREADS) {
        if (ci[k] == i) {
            Fr v = fr_zero();
            for (int j = 0; j < 4; ++j) v.v[j] = r[4 * k + j];
            acc = fr_add(acc, v);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = KZ_COL_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fr_add(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int j = 0; j < 8; ++j) out[8 * i + j] = red[0].v[j];
}
void launch_cell_reductions(cudaStream_t s, const Fr* coefs, const uint32_t* ci, const uint32_t* r, size_t m, uint32_t nc,
                            uint32_t* w_out /*8 nc*/, uint32_t* negS_out /*8*64*/) {
    k_cell_column_sum<<<KZ_CELL_LEN, KZ_COL_THREADS, 0, s>>>(coefs, m, negS_out);
    KZ_COUNT_LAUNCH();
    k_cell_commit_weights<<<nc, KZ_COL_THREADS, 0, s>>>(ci, r, m, w_out);
    KZ_COUNT_LAUNCH();
}
