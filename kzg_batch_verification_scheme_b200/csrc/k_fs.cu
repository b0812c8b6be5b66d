// K2/K3: batched Fiat-Shamir (BASELINE.json:5 item (c)) -- SHA-256 leaf and chunk hashes, challenge
// derivation r_i, the scalar products r_i z_i and sum r_i y_i -- plus the device-side synthetic instance
// generator and the host root hash.
#include "kernels.h"
#include "eip4844.cuh"
#include "sha256.cuh"

__device__ __forceinline__ u32 ld_be32(u32 x) { return __byte_perm(x, 0, 0x0123); }

// ---- leaf hashes: one thread per proof; also counts z_i, y_i >= r
__global__ void __launch_bounds__(128) k_leaf_hash(const u8* __restrict__ C, const u8* __restrict__ z,
                                                   const u8* __restrict__ y, const u8* __restrict__ pi, size_t n,
                                                   u32* __restrict__ leaves, u32* __restrict__ counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 cw[12], pw[12], zw[8], yw[8];
    const uint4* c4 = reinterpret_cast<const uint4*>(C + 48 * i);
    const uint4* p4 = reinterpret_cast<const uint4*>(pi + 48 * i);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint4 a = __ldg(c4 + k), b = __ldg(p4 + k);
        cw[4 * k] = ld_be32(a.x); cw[4 * k + 1] = ld_be32(a.y); cw[4 * k + 2] = ld_be32(a.z); cw[4 * k + 3] = ld_be32(a.w);
        pw[4 * k] = ld_be32(b.x); pw[4 * k + 1] = ld_be32(b.y); pw[4 * k + 2] = ld_be32(b.z); pw[4 * k + 3] = ld_be32(b.w);
    }
    const uint4* z4 = reinterpret_cast<const uint4*>(z + 32 * i);
    const uint4* y4 = reinterpret_cast<const uint4*>(y + 32 * i);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        uint4 a = __ldg(z4 + k), b = __ldg(y4 + k);
        zw[4 * k] = ld_be32(a.x); zw[4 * k + 1] = ld_be32(a.y); zw[4 * k + 2] = ld_be32(a.z); zw[4 * k + 3] = ld_be32(a.w);
        yw[4 * k] = ld_be32(b.x); yw[4 * k + 1] = ld_be32(b.y); yw[4 * k + 2] = ld_be32(b.z); yw[4 * k + 3] = ld_be32(b.w);
    }
    Fr zr, yr;
#pragma unroll
    for (int k = 0; k < 8; ++k) { zr.v[k] = zw[7 - k]; yr.v[k] = yw[7 - k]; }
    u32 bad = (fr_raw_is_canonical(zr) ? 0u : 1u) + (fr_raw_is_canonical(yr) ? 0u : 1u);
    if (bad) atomicAdd(counters + 1, bad);
    u32 h[8];
    fs_leaf_words(h, cw, zw, yw, pw);
    uint4* o4 = reinterpret_cast<uint4*>(leaves + 8 * i);
    o4[0] = make_uint4(h[0], h[1], h[2], h[3]);
    o4[1] = make_uint4(h[4], h[5], h[6], h[7]);
}
void launch_leaf_hash(cudaStream_t s, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                      uint32_t* leaves, uint32_t* counters) {
    if (!n) return;
    k_leaf_hash<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(C, z, y, pi, n, leaves, counters);
    KZ_COUNT_LAUNCH();
}

// ---- chunk digests: one thread per chunk of KZ_FS_CHUNK leaves (serial SHA-256 over <= 4 KiB)
__global__ void k_chunk_hash(const u32* __restrict__ leaves, size_t n, u32* __restrict__ digests) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nch = (n + KZ_FS_CHUNK - 1) / KZ_FS_CHUNK;
    if (j >= nch) return;
    size_t lo = j * KZ_FS_CHUNK;
    u32 cnt = (u32)((n - lo) < KZ_FS_CHUNK ? (n - lo) : KZ_FS_CHUNK);
    u32 h[8];
    fs_chunk_words(h, leaves + 8 * lo, cnt);
#pragma unroll
    for (int k = 0; k < 8; ++k) digests[8 * j + k] = h[k];
}
void launch_chunk_hash(cudaStream_t s, const uint32_t* leaves, size_t n, uint32_t* digests_words) {
    if (!n) return;
    size_t nch = (n + KZ_FS_CHUNK - 1) / KZ_FS_CHUNK;
    k_chunk_hash<<<(unsigned)((nch + 31) / 32), 32, 0, s>>>(leaves, n, digests_words);
    KZ_COUNT_LAUNCH();
}

// ---- challenges and scalar products
#define KZ_CH_THREADS 128
__global__ void __launch_bounds__(KZ_CH_THREADS) k_challenges(const u32* __restrict__ root_words, u64 global_offset,
                                                              const u8* __restrict__ z, const u8* __restrict__ y, size_t n,
                                                              int single, u32* __restrict__ r_out, u32* __restrict__ rz_out,
                                                              u32* __restrict__ partials) {
    __shared__ Fr red[KZ_CH_THREADS];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fr ry = fr_zero();
    if (i < n) {
        u32 root[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) root[k] = root_words[k];
        Fr r = fr_zero();
        if (single) r.v[0] = 1; else fs_r_limbs(r.v, root, global_offset + i);
        Fr zr, yr;
        const uint4* z4 = reinterpret_cast<const uint4*>(z + 32 * i);
        const uint4* y4 = reinterpret_cast<const uint4*>(y + 32 * i);
        uint4 a0 = __ldg(z4), a1 = __ldg(z4 + 1), b0 = __ldg(y4), b1 = __ldg(y4 + 1);
        u32 zw[8] = {ld_be32(a0.x), ld_be32(a0.y), ld_be32(a0.z), ld_be32(a0.w), ld_be32(a1.x), ld_be32(a1.y), ld_be32(a1.z), ld_be32(a1.w)};
        u32 yw[8] = {ld_be32(b0.x), ld_be32(b0.y), ld_be32(b0.z), ld_be32(b0.w), ld_be32(b1.x), ld_be32(b1.y), ld_be32(b1.z), ld_be32(b1.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) { zr.v[k] = zw[7 - k]; yr.v[k] = yw[7 - k]; }
        // raw * Montgomery -> canonical product
        Fr rz = fr_mul(r, fr_to_mont(zr));
        ry = fr_mul(r, fr_to_mont(yr));
        uint4* ro = reinterpret_cast<uint4*>(r_out + 4 * i);
        ro[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        uint4* zo = reinterpret_cast<uint4*>(rz_out + 8 * i);
        zo[0] = make_uint4(rz.v[0], rz.v[1], rz.v[2], rz.v[3]);
        zo[1] = make_uint4(rz.v[4], rz.v[5], rz.v[6], rz.v[7]);
    }
    red[threadIdx.x] = ry;
    __syncthreads();
    for (int s = KZ_CH_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fr_add(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) partials[8 * blockIdx.x + k] = red[0].v[k];
    }
}
// single block: sum the per-block partials; write sum (canonical) and -sum into scalar slot n of rz
__global__ void __launch_bounds__(KZ_CH_THREADS) k_sum_ry(const u32* __restrict__ partials, size_t nparts, size_t n,
                                                          u32* __restrict__ rz_out, u32* __restrict__ sum_out) {
    __shared__ Fr red[KZ_CH_THREADS];
    Fr acc = fr_zero();
    for (size_t j = threadIdx.x; j < nparts; j += KZ_CH_THREADS) {
        Fr v;
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] = partials[8 * j + k];
        acc = fr_add(acc, v);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = KZ_CH_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fr_add(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Fr neg = fr_neg(red[0]);
#pragma unroll
        for (int k = 0; k < 8; ++k) { sum_out[k] = red[0].v[k]; rz_out[8 * n + k] = neg.v[k]; }
    }
}
void launch_challenges(cudaStream_t s, const uint32_t* root_words, uint64_t global_offset, const uint8_t* z,
                       const uint8_t* y, size_t n, int single, uint32_t* r_out, uint32_t* rz_out, uint32_t* partials,
                       uint32_t* sum_ry_out) {
    if (!n) return;
    size_t nb = (n + KZ_CH_THREADS - 1) / KZ_CH_THREADS;
    k_challenges<<<(unsigned)nb, KZ_CH_THREADS, 0, s>>>(root_words, global_offset, z, y, n, single, r_out, rz_out, partials);
    KZ_COUNT_LAUNCH();
    k_sum_ry<<<1, KZ_CH_THREADS, 0, s>>>(partials, nb, n, rz_out, sum_ry_out);
    KZ_COUNT_LAUNCH();
}
// ---- EIP-4844 transcript mode (eip4844.cuh): powers of one challenge
__global__ void k_eip_table(const u8* __restrict__ hash_be, Fr* __restrict__ table, u32* __restrict__ r_out) {
    if (threadIdx.x || blockIdx.x) return;
    Fr r;
    eip_power_table(table, r, hash_be);
    for (int k = 0; k < 8; ++k) r_out[k] = r.v[k];
}
// rpow[i] = r^i (8 limbs; slot n = 0), rz[i] = r^i z_i, block partials of sum r^i y_i; counters[1] += scalars >= r
__global__ void __launch_bounds__(KZ_CH_THREADS) k_eip_scalars(const Fr* __restrict__ table, const u8* __restrict__ z, const u8* __restrict__ y,
                                                               size_t n, u32* __restrict__ rpow_out, u32* __restrict__ rz_out,
                                                               u32* __restrict__ partials, u32* __restrict__ counters) {
    __shared__ Fr red[KZ_CH_THREADS];
    __shared__ Fr tab[KZ_EIP_POW_BITS];
    if (threadIdx.x < KZ_EIP_POW_BITS) tab[threadIdx.x] = table[threadIdx.x];
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fr ry = fr_zero();
    if (i < n) {
        Fr r = eip_power(tab, (u64)i), zr, yr;
        fr_raw_from_be(zr, z + 32 * i);
        fr_raw_from_be(yr, y + 32 * i);
        u32 bad = (fr_raw_is_canonical(zr) ? 0u : 1u) + (fr_raw_is_canonical(yr) ? 0u : 1u);
        if (bad) { atomicAdd(counters + 1, bad); zr = fr_zero(); yr = fr_zero(); }
        Fr rz = fr_mul(r, fr_to_mont(zr));
        ry = fr_mul(r, fr_to_mont(yr));
        for (int k = 0; k < 8; ++k) { rpow_out[8 * i + k] = r.v[k]; rz_out[8 * i + k] = rz.v[k]; }
    } else if (i == n) {
        for (int k = 0; k < 8; ++k) rpow_out[8 * i + k] = 0;          // scalar of the setup point in the sum over the proofs
    }
    red[threadIdx.x] = ry;
    __syncthreads();
    for (int s = KZ_CH_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fr_add(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 8; ++k) partials[8 * blockIdx.x + k] = red[0].v[k];
}
void launch_eip_scalars(cudaStream_t s, const uint8_t* hash_be_dev, Fr* table, const uint8_t* z, const uint8_t* y, size_t n,
                        uint32_t* r_out, uint32_t* rpow_out, uint32_t* rz_out, uint32_t* partials, uint32_t* sum_ry_out, uint32_t* counters) {
    if (!n) return;
    k_eip_table<<<1, 32, 0, s>>>(hash_be_dev, table, r_out);
    KZ_COUNT_LAUNCH();
    size_t nb = (n + 1 + KZ_CH_THREADS - 1) / KZ_CH_THREADS;
    k_eip_scalars<<<(unsigned)nb, KZ_CH_THREADS, 0, s>>>(table, z, y, n, rpow_out, rz_out, partials, counters);
    KZ_COUNT_LAUNCH();
    k_sum_ry<<<1, KZ_CH_THREADS, 0, s>>>(partials, nb, n, rz_out, sum_ry_out);
    KZ_COUNT_LAUNCH();
}
// m 32-byte big-endian hashes -> the same values mod r (blob challenges)
__global__ void k_eip_reduce_be(u8* __restrict__ io, size_t m) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fr v;
    fr_raw_from_be(v, io + 32 * i);
    v = fr_reduce_raw(v);
    fr_raw_to_be(io + 32 * i, v);
}
void launch_eip_reduce_be(cudaStream_t s, uint8_t* io, size_t m) {
    if (!m) return;
    k_eip_reduce_be<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(io, m);
    KZ_COUNT_LAUNCH();
}

// r_i only, as 16 big-endian bytes (stage export kzgb_fs_challenges)
__global__ void k_r_only(const u32* __restrict__ root_words, size_t n, u8* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 root[8], r[4];
    for (int k = 0; k < 8; ++k) root[k] = root_words[k];
    fs_r_limbs(r, root, i);
    for (int k = 0; k < 4; ++k) {
        u32 v = r[3 - k];
        out[16 * i + 4 * k] = v >> 24; out[16 * i + 4 * k + 1] = v >> 16; out[16 * i + 4 * k + 2] = v >> 8; out[16 * i + 4 * k + 3] = v;
    }
}
void launch_r_only(cudaStream_t s, const uint32_t* root_words, size_t n, uint8_t* r_be16) {
    if (!n) return;
    k_r_only<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(root_words, n, r_be16);
    KZ_COUNT_LAUNCH();
}
// 32-byte big-endian scalars -> 8 little-endian limbs (kzgb_g1_msm); counts values >= r
__global__ void k_scalars_from_be(const u8* __restrict__ in, size_t m, u32* __restrict__ out, u32* counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fr v;
    fr_raw_from_be(v, in + 32 * i);
    if (!fr_raw_is_canonical(v)) atomicAdd(counters + 1, 1u);
    for (int k = 0; k < 8; ++k) out[8 * i + k] = v.v[k];
}
void launch_scalars_from_be(cudaStream_t s, const uint8_t* be32, size_t m, uint32_t* limbs8, uint32_t* counters) {
    if (!m) return;
    k_scalars_from_be<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(be32, m, limbs8, counters);
    KZ_COUNT_LAUNCH();
}

// ---- synthetic instances on the device (scalar shortcut); not part of the timed path.
// comb table: tab[w*255 + j-1] = j * 2^(8w) * G1 (affine, Montgomery), w < 32, j = 1..255.
__global__ void k_build_comb(Fp* tab) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32) return;
    G1Aff g = {fp_const(G1_GEN_X), fp_const(G1_GEN_Y)};
    G1Jac base = jac_from_aff(g);
    for (int k = 0; k < 8 * w; ++k) base = jac_dbl(base);
    G1Jac acc = base;
    for (int j = 0; j < 255; ++j) {
        G1Aff a = jac_to_aff(acc);
        tab[2 * (w * 255 + j)] = a.x;
        tab[2 * (w * 255 + j) + 1] = a.y;
        acc = jac_add(acc, base);
    }
}
void launch_build_comb(cudaStream_t s, Fp* comb_table) {
    k_build_comb<<<1, 32, 0, s>>>(comb_table);
    KZ_COUNT_LAUNCH();
}
__device__ __noinline__ G1Aff comb_mul(const Fp* tab, const Fr& k_mont) {
    Fr k = fr_from_mont(k_mont);
    G1Jac acc = jac_inf();
    for (int w = 0; w < 32; ++w) {
        u32 d = (k.v[w >> 2] >> (8 * (w & 3))) & 0xFF;
        if (d) {
            G1Aff t = {tab[2 * (w * 255 + d - 1)], tab[2 * (w * 255 + d - 1) + 1]};
            acc = jac_madd(acc, t);
        }
    }
    return jac_to_aff(acc);
}
__device__ __forceinline__ void store_be_words(u8* dst, const u32* w, int nw) {
    u32* d = reinterpret_cast<u32*>(dst);
    for (int k = 0; k < nw; ++k) d[k] = __byte_perm(w[k], 0, 0x0123);
}
__global__ void __launch_bounds__(64) k_synth(u64 seed, u64 offset, size_t n, const Fp* __restrict__ tab, u8* C, u8* z,
                                              u8* y, u8* pi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 idx = offset + i;
    Fr a = prng_fr(seed, 1, idx), zz = prng_fr(seed, 2, idx), yy = prng_fr(seed, 3, idx);
    Fr den = fr_sub(fr_const(FR_TEST_TAU), zz);
    Fr q = fr_mul(fr_sub(a, yy), fr_inv(den));
    G1Aff cp = comb_mul(tab, a), pp = comb_mul(tab, q);
    u32 w[12];
    g1_compress_words(w, cp);
    store_be_words(C + 48 * i, w, 12);
    g1_compress_words(w, pp);
    store_be_words(pi + 48 * i, w, 12);
    Fr zc = fr_from_mont(zz), yc = fr_from_mont(yy);
    u32 t[8];
    for (int k = 0; k < 8; ++k) t[k] = zc.v[7 - k];
    store_be_words(z + 32 * i, t, 8);
    for (int k = 0; k < 8; ++k) t[k] = yc.v[7 - k];
    store_be_words(y + 32 * i, t, 8);
}
void launch_synth(cudaStream_t s, uint64_t seed, uint64_t offset, size_t n, const Fp* comb_table, uint8_t* C, uint8_t* z,
                  uint8_t* y, uint8_t* pi) {
    if (!n) return;
    k_synth<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(seed, offset, n, comb_table, C, z, y, pi);
    KZ_COUNT_LAUNCH();
}

// ---- host SHA-256 for the root (32 B per 128 proofs: 256 KiB at n = 2^20, 2 MiB for 8 x 2^20; the one cross-shard step).
// Whole blocks go straight from the input; with the x86 SHA extensions (checked with cpuid at run time) the
// compression runs on sha256rnds2 / sha256msg1 / sha256msg2, otherwise on the portable rounds.
#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif
namespace {
static const uint32_t HK[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

static inline uint32_t hrotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_blocks_portable(uint32_t h[8], const uint8_t* p, size_t nblocks) {
    for (; nblocks; --nblocks, p += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = hrotr(w[i - 15], 7) ^ hrotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = hrotr(w[i - 2], 17) ^ hrotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (hrotr(e, 6) ^ hrotr(e, 11) ^ hrotr(e, 25)) + ((e & f) ^ (~e & g)) + HK[i] + w[i];
            uint32_t t2 = (hrotr(a, 2) ^ hrotr(a, 13) ^ hrotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
}
#if defined(__x86_64__)
// SHA extensions: the state lives in two registers as (A,B,E,F) and (C,D,G,H); four rounds per sha256rnds2 pair
__attribute__((target("sha,sse4.1,ssse3"))) static void sha_blocks_ni(uint32_t h[8], const uint8_t* p, size_t nblocks) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bLL, 0x0405060700010203LL);
    __m128i tmp = _mm_loadu_si128((const __m128i*)&h[0]);        // DCBA
    __m128i st1 = _mm_loadu_si128((const __m128i*)&h[4]);        // HGFE
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                          // CDAB
    st1 = _mm_shuffle_epi32(st1, 0x1B);                          // EFGH
    __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                  // ABEF
    st1 = _mm_blend_epi16(st1, tmp, 0xF0);                       // CDGH
    for (; nblocks; --nblocks, p += 64) {
        const __m128i save0 = st0, save1 = st1;
        __m128i m[4];
        for (int i = 0; i < 4; ++i) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), bswap);
        for (int r = 0; r < 16; ++r) {                           // 16 groups of four rounds
            __m128i cur = m[r & 3];
            __m128i wk = _mm_add_epi32(cur, _mm_loadu_si128((const __m128i*)&HK[4 * r]));
            st1 = _mm_sha256rnds2_epu32(st1, st0, wk);
            st0 = _mm_sha256rnds2_epu32(st0, st1, _mm_shuffle_epi32(wk, 0x0E));
            if (r < 12) {
                // message schedule: w[t+16..t+19] from the four most recent groups (cur = w[t..t+3] is the oldest)
                __m128i w1 = m[(r + 1) & 3], w2 = m[(r + 2) & 3], w3 = m[(r + 3) & 3];
                __m128i x = _mm_sha256msg1_epu32(cur, w1);
                x = _mm_add_epi32(x, _mm_alignr_epi8(w3, w2, 4));
                m[r & 3] = _mm_sha256msg2_epu32(x, w3);
            }
        }
        st0 = _mm_add_epi32(st0, save0);
        st1 = _mm_add_epi32(st1, save1);
    }
    tmp = _mm_shuffle_epi32(st0, 0x1B);                          // FEBA
    st1 = _mm_shuffle_epi32(st1, 0xB1);                          // DCHG
    st0 = _mm_blend_epi16(tmp, st1, 0xF0);                       // DCBA
    st1 = _mm_alignr_epi8(st1, tmp, 8);                          // HGFE
    _mm_storeu_si128((__m128i*)&h[0], st0);
    _mm_storeu_si128((__m128i*)&h[4], st1);
}
static bool cpu_has_sha() {
    unsigned a = 0, b = 0, c = 0, d = 0;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
    bool sha = (b >> 29) & 1u;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    return sha && ((c >> 19) & 1u) && ((c >> 9) & 1u);           // SHA, SSE4.1, SSSE3
}
#endif
static void sha_blocks(uint32_t h[8], const uint8_t* p, size_t nblocks) {
#if defined(__x86_64__)
    static const bool ni = cpu_has_sha() && !getenv("KZGB_NO_SHANI");
    if (ni) { sha_blocks_ni(h, p, nblocks); return; }
#endif
    sha_blocks_portable(h, p, nblocks);
}
struct HostSha {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    HostSha() {
        static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(h, iv, sizeof h);
    }
    void update(const uint8_t* p, size_t n) {
        size_t fill = (size_t)(len % 64);
        len += n;
        if (fill) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            p += take; n -= take;
            if (fill + take < 64) return;
            sha_blocks(h, buf, 1);
        }
        if (n >= 64) { sha_blocks(h, p, n / 64); p += n & ~(size_t)63; n &= 63; }
        if (n) memcpy(buf, p, n);
    }
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad[72] = {0x80};
        size_t fill = (size_t)(len % 64), padlen = (fill < 56 ? 56 : 120) - fill;
        update(pad, padlen);
        uint8_t lb[8];
        for (int i = 0; i < 8; ++i) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; ++i) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
    }
};
}  // namespace
void host_sha256_root(uint8_t out[32], const uint8_t* digests, size_t n_chunks, uint64_t n_total) {
    HostSha s;
    s.update((const uint8_t*)"KZGB200/root_v1_", 16);
    uint8_t be[16];
    uint64_t deg = 4096;
    for (int i = 0; i < 8; ++i) { be[i] = (uint8_t)(deg >> (56 - 8 * i)); be[8 + i] = (uint8_t)(n_total >> (56 - 8 * i)); }
    s.update(be, 16);
    s.update(digests, 32 * n_chunks);
    s.final(out);
}

// EIP-4844: SHA256("RCKZGBATCH___V1_" | u64be(4096) | u64be(n) | C_0 | z_0 | y_0 | pi_0 | C_1 | ...) over SoA inputs
void host_eip4844_batch_hash(uint8_t out[32], const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n) {
    HostSha s;
    s.update((const uint8_t*)"RCKZGBATCH___V1_", 16);
    uint8_t be[16];
    const uint64_t deg = 4096, nn = n;
    for (int i = 0; i < 8; ++i) { be[i] = (uint8_t)(deg >> (56 - 8 * i)); be[8 + i] = (uint8_t)(nn >> (56 - 8 * i)); }
    s.update(be, 16);
    uint8_t rec[160];
    for (size_t i = 0; i < n; ++i) {
        memcpy(rec, C + 48 * i, 48); memcpy(rec + 48, z + 32 * i, 32); memcpy(rec + 80, y + 32 * i, 32); memcpy(rec + 112, pi + 48 * i, 48);
        s.update(rec, 160);
    }
    s.final(out);
}
// EIP-4844 compute_challenge: SHA256("FSBLOBVERIFY_V1_" | u128be(4096) | blob | commitment)
void host_eip4844_blob_hash(uint8_t out[32], const uint8_t* blob, const uint8_t* commitment) {
    HostSha s;
    s.update((const uint8_t*)"FSBLOBVERIFY_V1_", 16);
    uint8_t be[16] = {0};
    be[14] = 0x10;                                      // 4096 as a 16-byte big-endian integer
    s.update(be, 16);
    s.update(blob, 131072);
    s.update(commitment, 48);
    s.final(out);
}

// cell batch root: SHA256("KZGB200/croot_v1" | u64be(nc) | u64be(m) | SHA256("KZGB200/comm_v1_" | commitments) | chunk digests)
void host_sha256_cell_root(uint8_t out[32], const uint8_t* comms, size_t nc, const uint8_t* digests, size_t n_chunks, uint64_t m) {
    uint8_t cdig[32];
    HostSha c;
    c.update((const uint8_t*)"KZGB200/comm_v1_", 16);
    c.update(comms, 48 * nc);
    c.final(cdig);
    HostSha s;
    s.update((const uint8_t*)"KZGB200/croot_v1", 16);
    uint8_t be[16];
    uint64_t a = nc;
    for (int i = 0; i < 8; ++i) { be[i] = (uint8_t)(a >> (56 - 8 * i)); be[8 + i] = (uint8_t)(m >> (56 - 8 * i)); }
    s.update(be, 16);
    s.update(cdig, 32);
    s.update(digests, 32 * n_chunks);
    s.final(out);
}
