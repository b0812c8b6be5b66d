// Launcher prototypes shared by the translation units of libkzgb200.so (host side only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>

#include "msm.cuh"
#include "pairing.cuh"
#include "mpair.cuh"

extern std::atomic<uint64_t> g_kzgb_launches;      // counts every kernel launch of this library
#define KZ_COUNT_LAUNCH() (g_kzgb_launches.fetch_add(1, std::memory_order_relaxed))

// ---- k_decompress.cu
// points [0,n): from inC, [n,2n): from inPi (48 B compressed each).  out_pts: 2 Fp per point (Montgomery
// affine, (0,0) = infinity/invalid).  counters[0] += number of points with status != 0.
// tmp: 3 Fp per point of scratch (Jacobian [|x|]P between the two subgroup-check kernels).
void launch_decompress(cudaStream_t s, const uint8_t* inC, const uint8_t* inPi, size_t n, Fp* out_pts, Fp* tmp, uint8_t* status,
                       uint32_t* counters);
// k_subgroup.cu
void launch_subgroup_chains(cudaStream_t s, Fp* pts, size_t m, Fp* tmp, uint8_t* status, uint32_t* counters, int mb, int mc);
void launch_decompress_points(cudaStream_t s, const uint8_t* in, size_t m, Fp* out_pts, Fp* tmp, uint8_t* status,
                              uint32_t* counters);
// K1a only / per-point subgroup check only (batched subgroup check and its fallback)
void launch_decompress_sqrt_points(cudaStream_t s, const uint8_t* in, size_t m, Fp* out_pts, uint8_t* status, uint32_t* counters);
void launch_decompress_sqrt(cudaStream_t s, const uint8_t* inC, const uint8_t* inPi, size_t n, Fp* out_pts, uint8_t* status,
                            uint32_t* counters);
void launch_subgroup_points(cudaStream_t s, Fp* pts, size_t m, Fp* tmp, uint8_t* status, uint32_t* counters);
void launch_points_to_be(cudaStream_t s, const Fp* pts, size_t m, uint8_t* out96);           // canonical x||y
void launch_points_from_be(cudaStream_t s, const uint8_t* in96, size_t m, Fp* pts, uint32_t* counters);
void launch_debug_op(cudaStream_t s, int op, const uint8_t* in, uint8_t* out, size_t count);
void launch_imad_bench(cudaStream_t s, uint32_t* sink, int blocks, int threads, int iters, int mode);

// ---- k_fs.cu
void launch_leaf_hash(cudaStream_t s, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                      uint32_t* leaves, uint32_t* counters /* [1] += scalars >= r */);
void launch_chunk_hash(cudaStream_t s, const uint32_t* leaves, size_t n, uint32_t* digests_words);
// r_i (4 limbs), rz_i = r_i z_i (8 limbs), block partials of sum r_i y_i; then sum + slot n of rz = -sum
void launch_challenges(cudaStream_t s, const uint32_t* root_words, uint64_t global_offset, const uint8_t* z,
                       const uint8_t* y, size_t n, int single, uint32_t* r_out, uint32_t* rz_out, uint32_t* partials,
                       uint32_t* sum_ry_out /*8 limbs canonical*/);
void launch_r_only(cudaStream_t s, const uint32_t* root_words, size_t n, uint8_t* r_be16);
void launch_synth(cudaStream_t s, uint64_t seed, uint64_t offset, size_t n, const Fp* comb_table, uint8_t* C, uint8_t* z,
                  uint8_t* y, uint8_t* pi);
void launch_build_comb(cudaStream_t s, Fp* comb_table /* 32*255 affine points */);
void launch_scalars_from_be(cudaStream_t s, const uint8_t* be32, size_t m, uint32_t* limbs8, uint32_t* counters);
void host_sha256_root(uint8_t out[32], const uint8_t* digests, size_t n_chunks, uint64_t n_total);
// EIP-4844 transcript mode (eip4844.cuh): table = 32 Fr of scratch; r_out / sum_ry_out 8 limbs; rpow_out 8(n+1), rz_out 8(n+1) limbs
void launch_eip_scalars(cudaStream_t s, const uint8_t* hash_be_dev, Fr* table, const uint8_t* z, const uint8_t* y, size_t n,
                        uint32_t* r_out, uint32_t* rpow_out, uint32_t* rz_out, uint32_t* partials, uint32_t* sum_ry_out, uint32_t* counters);
void launch_eip_reduce_be(cudaStream_t s, uint8_t* io, size_t m);
void host_eip4844_batch_hash(uint8_t out[32], const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n);
void host_eip4844_blob_hash(uint8_t out[32], const uint8_t* blob, const uint8_t* commitment);

// ---- k_msm.cu
struct MsmWorkspace {
    uint32_t *keys, *vals, *keys_alt, *vals_alt;   // capacity entries each
    size_t capacity;
    uint32_t* bucket_start;                        // total_buckets + 2: exclusive scan of the key histogram
    uint32_t *count, *cursor;                      // total_buckets + 2 each: histogram, scatter cursors
    uint32_t* tile_sum;                            // 1024: tile sums of the bucket scan
    G1Xyzz* buckets;                               // max total buckets
    G1Xyzz* sg_work;                               // sg_work_entries(plan): run sums and row / column totals
    G1Xyzz* slices;                                // 256: slice sums of the last reduction
    G1Xyzz* winsums;                               // KZ_MSM_MAX_WINDOWS
    ChunkRecs recs;                                // partial records, capacity/4 + 1 chunks
    size_t max_buckets;
};
// GLV: 255-bit scalars (8 limbs) -> k1 (m x 4 limbs) | k2 (m x 4 limbs); points P -> phi(P) = (beta^2 x, y)
void launch_glv_split(cudaStream_t s, const uint32_t* scalars8, size_t m, uint32_t* out4);
void launch_endo_points(cudaStream_t s, const Fp* src, size_t m, Fp* dst);
uint32_t msm_chunk_len(size_t N);
// digits + histogram + scan (= bucket boundaries) + scatter for `m` scalars of `nl` limbs each
void msm_sort_stage(cudaStream_t s, const MsmPlan& plan, const uint32_t* scalars, int nl, size_t m, MsmWorkspace& ws);
// accumulate + reduce + combine over points `pts` (2 Fp each, m points) using the sorted entries in ws
void msm_accumulate_stage(cudaStream_t s, const MsmPlan& plan, const Fp* pts, size_t m, MsmWorkspace& ws);
void msm_reduce_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws, G1Jac* out);
void msm_window_sums_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws, bool want_all);   // slices + per-window totals
void msm_slices_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws, bool want_all);        // ws.slices only
void msm_winsums_stage(cudaStream_t s, const MsmPlan& plan, MsmWorkspace& ws);                      // ws.winsums from ws.slices
// batched subgroup check on the slice sums of two sums (after msm_window_sums_stage with want_all): counters[2] += sums outside G1
void launch_sg_check(cudaStream_t s, const MsmPlan& plan, const MsmWorkspace& wa, const MsmWorkspace& wb, uint32_t* counters,
                     int nsums = 2);      // nsums = 1: only wa
void msm_combine_stage(cudaStream_t s, const MsmPlan* const* plans, MsmWorkspace* const* wss, G1Jac* const* outs, int njobs);
// up to 3 jobs; each job needs its own buckets / sg_work / slices / winsums; the serial Horner chains run concurrently
void msm_reduce_stage_multi(cudaStream_t s, const MsmPlan* const* plans, MsmWorkspace* const* wss, G1Jac* const* outs, int njobs);

// ---- k_pairing.cu
// setup: decompress the two G2 points, subgroup-check, precompute lines.  status[0] = 1 on success.
void launch_g2_setup(cudaStream_t s, const uint8_t* g2_bytes /*192*/, G2Lines* lines /*2*/, int* status);
void launch_g1_setup(cudaStream_t s, const uint8_t* g1_bytes /*48*/, Fp* g1_pt /*2 Fp*/, int* status);
// partial = (S1 + S2') | S3 | sum_ry as canonical big-endian bytes (320 B)
void launch_make_partial(cudaStream_t s, const G1Jac* s1, const G1Jac* s2, const G1Jac* s3, const uint32_t* sum_ry,
                         uint8_t* partial_out);
// combine G partials (device copy of 320*G bytes) -> A, B (Jacobian) and run the pairing check
void launch_combine(cudaStream_t s, const uint8_t* partials, int n_partials, G1Jac* AB /*2*/, uint32_t* sum_ry_total);
void launch_pairing(cudaStream_t s, const G2Lines* lines, const G1Jac* AB, int* result);
void launch_points_jac_from_be(cudaStream_t s, const uint8_t* in96, int m, G1Jac* out);
// artefacts: S1,S2,S3,A,B affine canonical from the Jacobian sums (S2 = S2' + sum_ry * G)
void launch_artifacts(cudaStream_t s, const G1Jac* s1, const G1Jac* s2p, const G1Jac* s3, const uint32_t* sum_ry,
                      const Fp* g1_pt, uint8_t* out /*5*96 + 32*/);
void launch_set_ab(cudaStream_t s, const G1Jac* a, const G1Jac* b, G1Jac* AB);
void launch_fr_to_be(cudaStream_t s, const uint32_t* in, uint8_t* out32);
void launch_fr_sum(cudaStream_t s, const uint32_t* in, int m, uint32_t* out);       // out = sum of m canonical Fr (8 limbs each)
void launch_jac_to_affine_be(cudaStream_t s, const G1Jac* in, int m, uint8_t* out96);   // m <= 32
void launch_pairing_debug(cudaStream_t s, int op, const G2Lines* lines, const uint8_t* in, uint8_t* out);

// ---- k_mpair.cu (Horner-free pairing check, mpair.cuh)
// one-time: chain[2][33] = [2^(4t)] of the two G2 points in g2_bytes (2 x 96 B compressed), tab[66] = their Miller lines;
// status[0..1] = chain ok, status[2] = lines ok
void launch_mp_setup(cudaStream_t s, const uint8_t* g2_bytes, G2Aff* chain, G2Lines* tab, int* status);
// terms[0..33) = V_t(s1) + V_t(s2), terms[33..66) = -V_t(s3)
void launch_mp_terms(cudaStream_t s, const MpSumDesc& s1, const MpSumDesc& s2, const MpSumDesc& s3, G1Xyzz* terms);
// coef[66] from the terms of n_shards shards (terms_in[g * 66 + p])
void launch_mp_coefs(cudaStream_t s, const G1Xyzz* terms_in, int n_shards, MpCoef* coef);
// line products (all SMs), merge, serial check; part: mp_part_entries() Fp12, F: 63 Fp12; *result = 1 iff the product is one
void launch_mp_check(cudaStream_t s, const G2Lines* tab, const MpCoef* coef, Fp12* part, Fp12* F, int* result);
size_t mp_part_entries();
// artefacts: AB[0] = A, AB[1] = B (Jacobian) by Horner over the shards' terms
void launch_mp_ab(cudaStream_t s, const G1Xyzz* terms_in, int n_shards, G1Jac* AB);
// wire format of the shard-level ABI (KZGB_TERMS_BYTES per shard): canonical big-endian coordinates + sum r_i y_i
void launch_mp_terms_to_wire(cudaStream_t s, const G1Xyzz* terms, const uint32_t* sum_ry, uint8_t* out);
void launch_mp_terms_from_wire(cudaStream_t s, const uint8_t* in, int n_shards, G1Xyzz* terms, uint32_t* sum_ry_total, uint32_t* bad);

// ---- k_cells.cu (cell batch, BASELINE.json config[4])
void launch_cell_twiddles(cudaStream_t s, Fr* W /*8192*/);
// ---- k_blob.cu (blob batch): leaves = 128 x 8 words per blob of scratch; z_out / y_out 32 B big-endian per blob
void launch_blob_challenges(cudaStream_t s, const uint8_t* blobs, const uint8_t* comms, size_t m, uint32_t* leaves, uint8_t* z_out);
void launch_blob_eval(cudaStream_t s, const Fr* W, const uint8_t* blobs, const uint8_t* z_in, size_t m, uint8_t* y_out, uint32_t* counter);
void launch_cell_leaf_hash(cudaStream_t s, const uint32_t* ci, const uint32_t* xi, const uint8_t* cells, const uint8_t* proofs, size_t m,
                           uint32_t* leaves);
void launch_cell_scalars(cudaStream_t s, const Fr* W, const uint32_t* root_words, const uint32_t* ci, const uint32_t* xi, uint32_t nc,
                         const uint8_t* cells, size_t m, uint64_t k0, Fr* coefs, uint32_t* r_out, uint32_t* rh_out, uint32_t* counters);
void launch_cell_reductions(cudaStream_t s, const Fr* coefs, const uint32_t* ci, const uint32_t* r, size_t m, uint32_t nc,
                            uint32_t* w_out, uint32_t* negS_out);
void host_sha256_cell_root(uint8_t out[32], const uint8_t* comms, size_t nc, const uint8_t* digests, size_t n_chunks, uint64_t m);
