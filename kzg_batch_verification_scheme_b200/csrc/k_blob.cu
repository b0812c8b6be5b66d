// Blob batch kernels: per-blob challenge hashing and barycentric evaluation.
#include "blob.cuh"
#include "kernels.h"

// one thread per (blob, 1 KiB leaf)
__global__ void __launch_bounds__(128) k_blob_leaves(const u8* __restrict__ blobs, size_t m, u32* __restrict__ leaves) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * KZ_BLOB_LEAVES) return;
    u32 h[8];
    blob_leaf_words(h, reinterpret_cast<const u32*>(blobs + 1024 * t));
    for (int i = 0; i < 8; ++i) leaves[8 * t + i] = h[i];
}
// one thread per blob: z (32 big-endian bytes)
__global__ void __launch_bounds__(64) k_blob_z(const u8* __restrict__ comms, const u32* __restrict__ leaves, size_t m, u8* __restrict__ z_out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    Fr z = blob_z(reinterpret_cast<const u32*>(comms + 48 * j), leaves + 8 * KZ_BLOB_LEAVES * j);
    fr_raw_to_be(z_out + 32 * j, z);
}
// one block per blob: y = p(z); counter += blob elements >= r (or z >= r)
__global__ void __launch_bounds__(KZ_BLOB_THREADS) k_blob_eval(const Fr* __restrict__ W, const u8* __restrict__ blobs, const u8* __restrict__ z_in,
                                                               size_t m, u8* __restrict__ y_out, u32* __restrict__ counter) {
    __shared__ Fr red[KZ_BLOB_THREADS];
    __shared__ Fr hit;
    __shared__ u32 has_hit, bad;
    const size_t j = blockIdx.x;
    if (j >= m) return;
    if (threadIdx.x == 0) { has_hit = 0; bad = 0; }
    __syncthreads();
    Fr zr;
    fr_raw_from_be(zr, z_in + 32 * j);
    const bool z_ok = fr_raw_is_canonical(zr);
    const Fr z = fr_to_mont(z_ok ? zr : fr_zero());
    BlobLane L = blob_eval_lane(W, blobs + (size_t)KZ_BLOB_LEN * 32 * j, z, threadIdx.x);
    red[threadIdx.x] = L.sum;
    if (L.has_hit) { hit = L.hit; has_hit = 1; }             // at most one lane: the domain points are distinct
    if (L.bad) atomicAdd(&bad, L.bad);
    __syncthreads();
    for (int st = KZ_BLOB_THREADS / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) red[threadIdx.x] = fr_add(red[threadIdx.x], red[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Fr y = has_hit ? hit : blob_eval_finish(z, red[0]);
        fr_raw_to_be(y_out + 32 * j, fr_from_mont(y));
        u32 nb = bad + (z_ok ? 0u : 1u);
        if (nb) atomicAdd(counter, nb);
    }
}

void launch_blob_challenges(cudaStream_t s, const uint8_t* blobs, const uint8_t* comms, size_t m, uint32_t* leaves, uint8_t* z_out) {
    if (!m) return;
    k_blob_leaves<<<(unsigned)((m * KZ_BLOB_LEAVES + 127) / 128), 128, 0, s>>>(blobs, m, leaves);
    KZ_COUNT_LAUNCH();
    k_blob_z<<<(unsigned)((m + 63) / 64), 64, 0, s>>>(comms, leaves, m, z_out);
    KZ_COUNT_LAUNCH();
}
void launch_blob_eval(cudaStream_t s, const Fr* W, const uint8_t* blobs, const uint8_t* z_in, size_t m, uint8_t* y_out, uint32_t* counter) {
    if (!m) return;
    k_blob_eval<<<(unsigned)m, KZ_BLOB_THREADS, 0, s>>>(W, blobs, z_in, m, y_out, counter);
    KZ_COUNT_LAUNCH();
}
