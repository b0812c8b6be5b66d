/* kzgb200.h -- C ABI of the B200-native KZG batch verifier (BLS12-381).
 *
 * Drop-in boundary.  The upstream reference (KoonMing/KZG-Batch-Verification-Scheme, mounted at
 * /root/reference) contains only LICENSE:1-201 -- it has NO FFI, plugin or operator interface
 * that these entry points could replace.  The boundary is therefore the
 * one BASELINE.json:5 names: "a C++ host library behind a thin C ABI exposing
 * verify_kzg_proof / verify_kzg_proof_batch".  Each declaration below cites the spec text it
 * implements instead of a reference file:line (none exists).
 *
 * Two shared libraries export this exact symbol set so one test harness drives both:
 *   libkzgb200.so        the product: CUDA sm_100a, no CPU fallback
 *   libkzgb_oracle.so    the CPU oracle (oracle/), test infrastructure + reported CPU baseline
 *
 * Byte formats: G1 compressed = 48 B ZCash format (bit7 compressed, bit6
 * infinity, bit5 y-sign); Fr = 32 B big-endian, must be < r; affine outputs = 96 B x||y
 * big-endian canonical, infinity or invalid = 96 zero bytes.
 */
#ifndef KZGB200_H
#define KZGB200_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { KZGB_OK = 0, KZGB_BADARGS = 1, KZGB_ERROR = 2, KZGB_MALLOC = 3 } kzgb_ret;

/* per-point status bytes of the decompression stage (first failure wins; App. B.2) */
enum { KZGB_ST_OK = 0, KZGB_ST_BAD_FLAGS = 1, KZGB_ST_X_GE_P = 2, KZGB_ST_NOT_ON_CURVE = 3, KZGB_ST_NOT_IN_G1 = 4 };

#define KZGB_CHUNK 128u           /* proofs per Fiat-Shamir chunk digest (128, not the 1024 of SURVEY App. B --
                                     the serial SHA-256 over a chunk is on the critical path of small batches) */
#define KZGB_PARTIAL_BYTES 320u   /* per-shard partial: A_shard = S1+S2-(sum r_i y_i)G1 (Jacobian X|Y|Z, 144 B) | S3 (144 B) | sum r_i y_i (32 B) */
#define KZGB_N_STAGES 10

typedef struct kzgb_ctx kzgb_ctx; /* opaque: trusted setup + per-GPU workspaces */

typedef struct {
    uint8_t S1[96], S2[96], S3[96]; /* sum r_i C_i ; sum r_i z_i pi_i ; sum r_i pi_i  (canonical affine) */
    uint8_t A[96], B[96];           /* pairing inputs: A = S1+S2-(sum r_i y_i)G1 ; B = -S3 */
    uint8_t sum_ry[32];             /* sum r_i y_i mod r, big-endian */
    uint8_t root[32];               /* Fiat-Shamir root */
    uint64_t n;                     /* batch size the artefacts belong to */
    uint32_t n_bad_points;          /* count of C_i/pi_i with status != OK */
    uint32_t n_bad_scalars;         /* count of z_i/y_i >= r */
    /* device ms per stage of the last call (0 where not applicable):
       0 h2d, 1 decompress+validate, 2 leaf+chunk hashes, 3 root (host), 4 challenges+scalars,
       5 msm digits+sort, 6 msm bucket accumulate + reduction to window totals, 7 Horner combine + partial,
       8 pairing, 9 total */
    float stage_ms[KZGB_N_STAGES];
} kzgb_artifacts;

/* ---- context.  BJ:5 "EIP-4844-style trusted setup": g1_monomial = [tau^i]G1 (48 B each, n1 >= 1),
 * g2_monomial = [tau^i]G2 (96 B each, n2 >= 2).  devices = CUDA ordinals (NULL => {0}); the oracle
 * library ignores devices.  n_max = largest per-device shard the workspaces are sized for. */
kzgb_ret kzgb_ctx_create(kzgb_ctx **out, const uint8_t *g1_monomial, size_t n1, const uint8_t *g2_monomial,
                         size_t n2, const int *devices, int n_devices, size_t n_max);
void kzgb_ctx_free(kzgb_ctx *ctx);

/* ---- the two entry points BJ:5 names.  Host pointers.  Malformed input => KZGB_BADARGS, *ok=false;
 * well-formed but wrong proof => KZGB_OK, *ok=false; CUDA failure => KZGB_ERROR. */
kzgb_ret verify_kzg_proof(bool *ok, const uint8_t C[48], const uint8_t z[32], const uint8_t y[32],
                          const uint8_t pi[48], kzgb_ctx *ctx);
kzgb_ret verify_kzg_proof_batch(bool *ok, const uint8_t *C /*48n*/, const uint8_t *z /*32n*/, const uint8_t *y /*32n*/,
                                const uint8_t *pi /*48n*/, size_t n, kzgb_ctx *ctx);

/* ---- cell batch (BASELINE.json config[4]): m multi-point openings on cosets of 64 points
 * (PeerDAS-shaped; conventions in DESIGN.md "Cell batch").  commitments: nc unique 48-byte G1; opening k refers to
 * commitments[commitment_indices[k]] and cell cell_indices[k] (< 128), carries 64 evaluations (32 B big-endian each,
 * < r) and one proof.  Needs a context created with n1 >= 64 ([tau^j]G1) and n2 >= 65 ([tau^64]G2).  A context over several
 * devices shards the openings over them (contiguous ranges of at least 1024 openings, multiples of KZGB_CHUNK); the result
 * does not depend on the device count. */
kzgb_ret verify_cell_kzg_proof_batch(bool *ok, const uint8_t *commitments /*48 nc*/, size_t nc,
                                     const uint32_t *commitment_indices /*m*/, const uint32_t *cell_indices /*m*/,
                                     const uint8_t *cells /*2048 m*/, const uint8_t *proofs /*48 m*/, size_t m, kzgb_ctx *ctx);

/* ---- pipelined form of verify_kzg_proof_batch: up to `depth` batches in flight on device 0 of the context, each on its
 * own workspace (depth x the memory of one) and host thread, so the latency-bound tail of one batch (bucket reduction,
 * pairing: ~1 ms) runs under the decompression kernel of the next.  kzgb_pipeline_init once (1 <= depth <= 8; changing
 * the depth needs every ticket collected); submit returns a ticket at once -- KZGB_BADARGS if `depth` batches are
 * already in flight; wait blocks until that batch is done and returns what verify_kzg_proof_batch would have.  The
 * input buffers (host or, inputs_on_device != 0, device pointers valid on the context's device) must stay untouched
 * until wait returns.  Every ticket is collected exactly once, in any order.  kzgb_last_artifacts does not cover
 * pipelined batches. */
kzgb_ret kzgb_pipeline_init(kzgb_ctx *ctx, int depth);
kzgb_ret verify_kzg_proof_batch_submit(uint64_t *ticket_out, const uint8_t *C, const uint8_t *z, const uint8_t *y,
                                       const uint8_t *pi, size_t n, int inputs_on_device, kzgb_ctx *ctx);
kzgb_ret verify_kzg_proof_batch_wait(bool *ok, uint64_t ticket, kzgb_ctx *ctx);

/* ---- same batch check with inputs already resident in device memory of ctx device 0 (used for the
 * device-resident throughput figure; `stream` is a cudaStream_t or NULL).  The four device arrays are read with
 * 128-bit loads: each base pointer must be 16-byte aligned (cudaMalloc'ed memory is), else KZGB_BADARGS -- this holds
 * for every entry point that takes device pointers (kzgb_shard_phase1 with inputs_on_device, the submit form).
 * Oracle: host pointers. */
kzgb_ret verify_kzg_proof_batch_device(bool *ok, const uint8_t *dC, const uint8_t *dz, const uint8_t *dy,
                                       const uint8_t *dpi, size_t n, kzgb_ctx *ctx, void *stream);

/* ---- shard-level entry points (BJ:5 multi-GPU: "the point set shards contiguously across 1/2/4/8
 * GPUs ... each GPU returns one partial G1 sum of 144 bytes ... combined on the host, no NCCL").
 * One process per GPU drives its own shard; only chunk digests, the root and the partials cross
 * process boundaries.  n_local must be a multiple of KZGB_CHUNK unless this is the last shard. */
kzgb_ret kzgb_shard_phase1(kzgb_ctx *ctx, int slot, const uint8_t *C, const uint8_t *z, const uint8_t *y,
                           const uint8_t *pi, size_t n_local, int inputs_on_device, void *stream,
                           uint8_t *chunk_digests_out /*32*ceil(n_local/KZGB_CHUNK)*/, uint32_t *n_bad_out);
kzgb_ret kzgb_fs_root(uint8_t root_out[32], const uint8_t *chunk_digests, size_t n_chunks, uint64_t n_total);
kzgb_ret kzgb_shard_phase2(kzgb_ctx *ctx, int slot, const uint8_t root[32], uint64_t global_offset, void *stream,
                           uint8_t partial_out[KZGB_PARTIAL_BYTES]);
kzgb_ret kzgb_combine_verify(kzgb_ctx *ctx, const uint8_t *partials /*320*G*/, int n_partials, bool *ok);
/* Same exchange without the serial Horner tail on any shard (DESIGN.md "Horner-free pairing check"): instead of one
 * 144-byte sum per side a shard returns 2 x 33 G1 terms V_t with A_shard = sum_t 2^(4t) V_t and B_shard =
 * sum_t 2^(4t) V_(33+t); the combining side adds the shards' terms and pairs them with the fixed multiples
 * [2^(4t)]G2, [2^(4t)][tau]G2.  Record = 66 x (X | Y | ZZ | ZZZ, 48 B big-endian canonical each; the point is
 * (X/ZZ, Y/ZZZ), ZZ = 0 = infinity) | sum r_i y_i of the shard (32 B big-endian): KZGB_TERMS_BYTES.
 * kzgb_shard_phase2_terms returns once the record is on the host; kzgb_shard_finish then reports the shard's input
 * validation (KZGB_BADARGS and the counts if a point or scalar of the shard is malformed) -- the caller must
 * collect it from every shard before trusting the verdict of kzgb_combine_verify_terms. */
#define KZGB_N_TERMS 66u
#define KZGB_TERMS_BYTES (KZGB_N_TERMS * 192u + 32u)
kzgb_ret kzgb_shard_phase2_terms(kzgb_ctx *ctx, int slot, const uint8_t root[32], uint64_t global_offset, void *stream,
                                 uint8_t terms_out[KZGB_TERMS_BYTES]);
kzgb_ret kzgb_shard_finish(kzgb_ctx *ctx, int slot, uint32_t *n_bad_points_out, uint32_t *n_bad_scalars_out);
kzgb_ret kzgb_combine_verify_terms(kzgb_ctx *ctx, const uint8_t *terms /*KZGB_TERMS_BYTES*G*/, int n_shards, bool *ok);

/* ---- stage exports: artefacts for bit-exact GPU-vs-oracle diffs (BJ:5 "bit-exact ... every canonical
 * affine MSM output and every decompressed point") */
kzgb_ret kzgb_g1_decompress_batch(uint8_t *affine_out /*96m*/, uint8_t *status_out /*m*/, const uint8_t *in /*48m*/,
                                  size_t m, kzgb_ctx *ctx);
kzgb_ret kzgb_fs_challenges(uint8_t root_out[32], uint8_t *r_out /*16n*/, const uint8_t *C, const uint8_t *z,
                            const uint8_t *y, const uint8_t *pi, size_t n, kzgb_ctx *ctx);
/* scalars: 32 B big-endian each, < r.  nbits: 255 or 128 (scalars must fit).  Points canonical affine
 * (96 zero bytes = infinity, skipped).  Returns BADARGS for points off the curve encoding range. */
kzgb_ret kzgb_g1_msm(uint8_t affine_out[96], const uint8_t *points_affine /*96m*/, const uint8_t *scalars /*32m*/,
                     size_t m, int nbits, kzgb_ctx *ctx);
/* device ms of the last kzgb_g1_msm call: [0] digits+sort [1] accumulate [2] reduce+combine [3] total */
kzgb_ret kzgb_g1_msm_times(float ms_out[4], kzgb_ctx *ctx);
kzgb_ret kzgb_pairing_check(bool *ok, const uint8_t A_affine[96], const uint8_t B_affine[96], kzgb_ctx *ctx);
kzgb_ret kzgb_last_artifacts(kzgb_ctx *ctx, kzgb_artifacts *out);

/* ---- synthetic instances (BJ:5 "Synthetic instances are generated from a known test tau").
 * Scalar-shortcut generator: proofs [offset, offset+n) of the stream `seed`.  Product library: runs on
 * the device and (out_on_device != 0) leaves the bytes in the given device buffers. */
kzgb_ret kzgb_synth_instance(kzgb_ctx *ctx, uint64_t seed, uint64_t offset, size_t n, uint8_t *C, uint8_t *z,
                             uint8_t *y, uint8_t *pi, int out_on_device);
/* (The insecure test setup itself -- [tau^i]G1, [tau^i]G2 from the documented tau -- needs G2 scalar multiplication, which
 * the verification path never uses: it is generated by the oracle library only, include/kzgb200_testing.h.) */

/* ---- primitive-level debug operator (tests only): applies op to `count` records of canonical
 * big-endian operands; see KZGB_OP_* for record layouts. */
enum {
    KZGB_OP_FP_MUL = 1,   /* in 96 (a|b)  out 48 */
    KZGB_OP_FP_SQR = 2,   /* in 48 out 48 */
    KZGB_OP_FP_ADD = 3,   /* in 96 out 48 */
    KZGB_OP_FP_SUB = 4,   /* in 96 out 48 */
    KZGB_OP_FP_INV = 5,   /* in 48 out 48 (inv(0)=0) */
    KZGB_OP_FP_SQRT_CAND = 6, /* in 48 out 48: a^((p+1)/4) */
    KZGB_OP_FR_MUL = 7,   /* in 64 (a|b) out 32 */
    KZGB_OP_FR_ADD = 8,   /* in 64 out 32 */
    KZGB_OP_G1_ADD = 9,   /* in 192 (P|Q affine) out 96 */
    KZGB_OP_G1_DBL = 10,  /* in 96 out 96 */
    KZGB_OP_G1_MUL = 11,  /* in 96+32 (P | k BE) out 96 */
    KZGB_OP_G1_MUL_XSQ = 12, /* in 96 out 96 : [x^2]P via the two |x| chains */
    KZGB_OP_FP12_MUL = 13, /* in 1152 (a|b: 6 x Fp2 coefficients of w^0..w^5, each c0|c1) out 576 */
    KZGB_OP_FP12_FROB1 = 14, /* in 576 out 576 */
    KZGB_OP_FP12_FROB2 = 15, /* in 576 out 576 */
    KZGB_OP_FP12_INV = 16,   /* in 576 out 576 */
    KZGB_OP_FINAL_EXP = 17,  /* in 576 out 576 : f^(3(p^12-1)/r) */
    KZGB_OP_MILLER_FE = 18,  /* in 192 (A|B affine) out 576 : final_exp(miller(A,G2) miller(B,[tau]G2)) */
    KZGB_OP_SHA256_64 = 19   /* in 64 out 32 : SHA-256 of a 64-byte message */
};
kzgb_ret kzgb_debug_op(kzgb_ctx *ctx, int op, const uint8_t *in, uint8_t *out, size_t count);

/* ---- measurement helpers (product library only; oracle returns KZGB_ERROR) */
/* integer multiply-add issue-rate microbenchmarks, thread-level operations per second over all SMs:
 * kzgb_imad_peak   carry-chained 32x32->64 multiply-adds (SASS IMAD.WIDE.U32.X), the instruction the
 *                  Montgomery products consist of -- the roofline denominator;
 * kzgb_imad32_peak plain 32-bit IMAD (the pipe's nominal issue rate, for context). */
kzgb_ret kzgb_imad_peak(kzgb_ctx *ctx, double *imad_per_sec_out, double *ms_out);
kzgb_ret kzgb_imad32_peak(kzgb_ctx *ctx, double *imad_per_sec_out, double *ms_out);
/* per-stage device ms of the last verify call without computing the artefacts (see kzgb_artifacts.stage_ms) */
kzgb_ret kzgb_last_stage_ms(kzgb_ctx *ctx, float ms_out[KZGB_N_STAGES]);
/* number of kernel launches issued by this library since ctx creation (for bench gpu_launches) */
uint64_t kzgb_launch_count(const kzgb_ctx *ctx);
/* threads the oracle uses (oracle library only; product returns 0) */
int kzgb_set_threads(kzgb_ctx *ctx, int n_threads);
/* ---- Blob batch (DESIGN.md "Blob batch"): the caller one step before the hot path, shaped
 * like c-kzg-4844's verify_blob_kzg_proof_batch(ok, blobs, commitments_bytes, proofs_bytes, n, settings).
 * blobs: m x 4096 field elements (32 B big-endian, < r), the evaluations of a polynomial of degree < 4096 over the
 * 4096-th roots of unity in bit-reversed order.  Per blob: z = hash(blob, commitment) mod r, y = p(z) (barycentric),
 * then the plain batch verify_kzg_proof_batch(C, z, y, proofs).  Hashing convention and domain are this library's
 * (not EIP-4844 wire compatible).  KZGB_BADARGS: a blob element >= r or a malformed point. */
kzgb_ret verify_blob_kzg_proof_batch(bool *ok, const uint8_t *blobs, const uint8_t *commitments, const uint8_t *proofs,
                                     size_t m, kzgb_ctx *ctx);
/* stage export: z_out, y_out (m x 32 B big-endian) of the blob batch above */
kzgb_ret kzgb_blob_challenges_evals(uint8_t *z_out, uint8_t *y_out, const uint8_t *blobs, const uint8_t *commitments,
                                    size_t m, kzgb_ctx *ctx);
/* stage export: y_out[j] = p_j(z_in[j]) for caller-chosen points (covers z on the evaluation domain) */
kzgb_ret kzgb_blob_eval(uint8_t *y_out, const uint8_t *blobs, const uint8_t *z_in, size_t m, kzgb_ctx *ctx);
#define KZGB_BLOB_BYTES 131072

/* ---- EIP-4844 / c-kzg-4844 transcript mode (DESIGN.md "EIP-4844 mode").  Same inputs, same
 * verdict semantics and return codes as verify_kzg_proof_batch, but the random linear combination is c-kzg's:
 *   r = int_be(SHA256("RCKZGBATCH___V1_" | u64be(4096) | u64be(n) | C_0 | z_0 | y_0 | pi_0 | C_1 | ...)) mod r_BLS,
 *   coefficients r^0 .. r^(n-1);  accept  <=>  e(sum r^i (C_i - y_i G1) + sum r^i z_i pi_i, G2) e(-sum r^i pi_i, [tau]G2) = 1.
 * The transcript hash is serial (host, SHA extensions: ~0.5 ms per 1000 proofs); every point gets the per-point
 * subgroup check.  Needs 2(n+1) <= n_max of the context; single device (slot 0).  kzgb_last_artifacts gives A, B,
 * sum_ry and, as `root`, the transcript digest.
 * Blob form: z_j = int_be(SHA256("FSBLOBVERIFY_V1_" | u128be(4096) | blob_j | C_j)) mod r_BLS, y_j = p_j(z_j) over the
 * bit-reversed 4096-th roots of unity of 7^((r-1)/4096), then the batch above. */
kzgb_ret verify_kzg_proof_batch_eip4844(bool *ok, const uint8_t *C, const uint8_t *z, const uint8_t *y, const uint8_t *pi,
                                        size_t n, kzgb_ctx *ctx);
kzgb_ret verify_blob_kzg_proof_batch_eip4844(bool *ok, const uint8_t *blobs, const uint8_t *commitments,
                                             const uint8_t *proofs, size_t m, kzgb_ctx *ctx);
kzgb_ret kzgb_blob_challenges_evals_eip4844(uint8_t *z_out, uint8_t *y_out, const uint8_t *blobs, const uint8_t *commitments,
                                            size_t m, kzgb_ctx *ctx);
/* Context from a c-kzg-4844 `trusted_setup.txt` (n_g1, n_g2, n_g1 G1 Lagrange points, n_g2 G2 monomials, optionally n_g1
 * G1 monomials; hex, one point per line).  KZGB_BADARGS: unreadable or malformed file, point not in its group. */
kzgb_ret kzgb_load_trusted_setup_file(kzgb_ctx **out, const char *path, const int *devices, int n_devices, size_t n_max);

/* Batches (shards) of at least n_min proofs establish subgroup membership of all 2n points through 128 slice
 * sums per MSM of the bucket tables that sum r_i C_i and sum r_i pi_i fill anyway (soundness 2^-128 per point,
 * DESIGN.md "Batched subgroup check"); smaller ones, and any batch in which a slice sum fails, run the
 * deterministic per-point check.  0 = always per point.  Default 2: every batch (env KZGB_SG_BATCH_MIN). */
kzgb_ret kzgb_set_subgroup_batch_min(kzgb_ctx *ctx, size_t n_min);
const char *kzgb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KZGB200_H */
