// Kernels of the Horner-free pairing check (mpair.cuh): one-time line tables for the 2 x 33 fixed G2 multiples, then
// per batch  terms -> coefficients -> line products (all SMs) -> merge -> serial check (one block).
#include "kernels.h"
#include "mpair.cuh"

// ------------------------------------------------------------------ one-time setup
// chain[b][t] = [2^(4t)] Q_b for the two G2 setup points (affine twist coordinates); status[b] = 1 on success
__global__ void k_mp_setup_chain(const u8* __restrict__ g2_bytes, G2Aff* chain, int* status) {
    const int b = blockIdx.x;
    if (threadIdx.x) return;
    if (b == 0) status[2] = 1;                       // cleared by k_mp_setup_lines on a degenerate step
    G2Aff q;
    bool ok = g2_decompress(q, g2_bytes + 96 * b);
    Fp2 la, lb;
    for (int t = 0; ok && t < KZ_MP_TERMS; ++t) {
        chain[b * KZ_MP_TERMS + t] = q;
        for (int u = 0; ok && u < KZ_MP_G && t + 1 < KZ_MP_TERMS; ++u) ok = g2_dbl_step(q, la, lb);
    }
    status[b] = ok ? 1 : 0;
}
// lines of the Miller loop for each of the 66 points; one working thread per block (data-dependent inversion loops)
__global__ void k_mp_setup_lines(const G2Aff* __restrict__ chain, G2Lines* tab, int* status) {
    if (threadIdx.x) return;
    G2Aff out;
    if (!g2_mul_xabs(out, chain[blockIdx.x], &tab[blockIdx.x])) atomicExch(status + 2, 0);
}
void launch_mp_setup(cudaStream_t s, const uint8_t* g2_bytes, G2Aff* chain, G2Lines* tab, int* status) {
    k_mp_setup_chain<<<2, 32, 0, s>>>(g2_bytes, chain, status);
    KZ_COUNT_LAUNCH();
    k_mp_setup_lines<<<KZ_MP_PAIRS, 32, 0, s>>>(chain, tab, status);
    KZ_COUNT_LAUNCH();
}

// ------------------------------------------------------------------ per batch
// terms[t] = V_t(S1) + V_t(S2'),  terms[33 + t] = -V_t(S3).  Two terms per block: quads (term, sum).
__global__ void __launch_bounds__(32) k_mp_terms(const MpSumDesc s1, const MpSumDesc s2, const MpSumDesc s3, G1Xyzz* __restrict__ terms) {
    __shared__ Fp qsm[8 * KZ_QUAD_SLOTS];
    __shared__ G1Xyzz ex[2];
    const int qi = threadIdx.x >> 2, tt = qi / 3, sum = qi - 3 * tt, t = blockIdx.x * 2 + tt;
    const bool active = qi < 6 && t < KZ_MP_TERMS;
    Quad q = quad_make(qsm, qi);
    G1Xyzz acc = xyzz_inf();
    if (active) {
        const MpSumDesc d = sum == 0 ? s1 : (sum == 1 ? s2 : s3);
        acc = mp_term(q, d, t);
    }
    if (active && sum == 1 && q.ql == 0) ex[tt] = acc;
    __syncthreads();
    if (active && sum == 0) {
        acc = quad_xyzz_add(q, acc, ex[tt]);
        if (q.ql == 0) terms[t] = acc;
    }
    if (active && sum == 2 && q.ql == 0) terms[KZ_MP_TERMS + t] = xyzz_neg(acc);
}
void launch_mp_terms(cudaStream_t s, const MpSumDesc& s1, const MpSumDesc& s2, const MpSumDesc& s3, G1Xyzz* terms) {
    k_mp_terms<<<(KZ_MP_TERMS + 1) / 2, 32, 0, s>>>(s1, s2, s3, terms);
    KZ_COUNT_LAUNCH();
}
// pair p: sum over the shards' terms, then the three products every line of the pair needs
__global__ void __launch_bounds__(32) k_mp_coefs(const G1Xyzz* __restrict__ terms_in, int n_shards, MpCoef* __restrict__ coef) {
    __shared__ Fp qsm[8 * KZ_QUAD_SLOTS];
    const int qi = threadIdx.x >> 2, p = blockIdx.x * 8 + qi;
    if (p >= KZ_MP_PAIRS) return;
    Quad q = quad_make(qsm, qi);
    G1Xyzz acc = terms_in[p];
    for (int g = 1; g < n_shards; ++g) acc = quad_xyzz_add(q, acc, terms_in[(size_t)g * KZ_MP_PAIRS + p]);
    const MpCoef c = mp_coef_of(q, acc);
    if (q.ql == 0) coef[p] = c;
}
void launch_mp_coefs(cudaStream_t s, const G1Xyzz* terms_in, int n_shards, MpCoef* coef) {
    k_mp_coefs<<<(KZ_MP_PAIRS + 7) / 8, 32, 0, s>>>(terms_in, n_shards, coef);
    KZ_COUNT_LAUNCH();
}

// Line products.  Block (step s, group g) evaluates the lines of step s at the pairs p = g, g + NG, ... and multiplies
// them with a tree of Fp12 products, MP_UNITS products per pass.  part[s * NG + g] = the group's product.
#define MP_UNITS 2
#define MP_NG 2
#define MP_POOL ((KZ_MP_PAIRS + MP_NG - 1) / MP_NG)
__global__ void __launch_bounds__(128 * MP_UNITS) k_mp_lines(const G2Lines* __restrict__ tab, const MpCoef* __restrict__ coef,
                                                              Fp12* __restrict__ part) {
    __shared__ Fp12 pool[MP_POOL];
    __shared__ MpUnit units[MP_UNITS];
    const int s = blockIdx.x, grp = blockIdx.y;
    const int unit = threadIdx.x >> 7, tu = threadIdx.x & 127;
    const int nl = (KZ_MP_PAIRS - grp + MP_NG - 1) / MP_NG;
    if (tu == 0) units[unit].zero = fp_zero();
    // line values as dense Fp12: l = a alpha + (b beta) w^2 + gamma w^3; a pair at infinity contributes 1
    for (int idx = threadIdx.x; idx < nl * 12; idx += blockDim.x) {
        const int l = idx / 12, ci = idx - 12 * l, p = grp + l * MP_NG;
        const Fp v = mp_line_coeff(tab[p], s, coef[p], ci);
        if (ci & 1) pool[l].c[ci >> 1].c1 = v; else pool[l].c[ci >> 1].c0 = v;
    }
    __syncthreads();
    int n = nl;
    while (n > 1) {
        const int m = n >> 1;
        for (int base = 0; base < m; base += MP_UNITS) {
            const int k = base + unit;
            if (k < m) mp_mul_products(units[unit], tu, pool[2 * k], pool[2 * k + 1]);
            __syncthreads();
            if (k < m) mp_mul_fold(units[unit], tu, pool[k]);
            __syncthreads();
        }
        if (n & 1) {
            if (threadIdx.x < 12) {
                const int k = threadIdx.x >> 1;
                if (threadIdx.x & 1) pool[m].c[k].c1 = pool[n - 1].c[k].c1; else pool[m].c[k].c0 = pool[n - 1].c[k].c0;
            }
            n = m + 1;
            __syncthreads();
        } else n = m;
    }
    if (threadIdx.x < 12) {
        const int k = threadIdx.x >> 1;
        Fp12& dst = part[s * MP_NG + grp];
        if (threadIdx.x & 1) dst.c[k].c1 = pool[0].c[k].c1; else dst.c[k].c0 = pool[0].c[k].c0;
    }
}
// Miller iterations are merged in chunks of MP_CHUNK: with F_it = product of the partial products of iteration it (its
// doubling step and, where bit 62 - it of |x| is set, the addition step that follows),
//     f <- f^(2^s) G_c,    G_c = ((F_a^2 F_(a+1))^2 ... )^2 F_(a+s-1)        (a = MP_CHUNK c, s = iterations in the chunk)
// so the serial kernel spends s squarings and ONE product per chunk instead of s of each; the G_c are independent of f
// and are formed here, one block per chunk.
#define MP_CHUNK 4
#define MP_NCHUNK ((KZ_MP_ITERS + MP_CHUNK - 1) / MP_CHUNK)
__global__ void __launch_bounds__(128) k_mp_merge(const Fp12* __restrict__ part, Fp12* __restrict__ G) {
    __shared__ MpUnit U;
    __shared__ Fp12 acc, fit, y;
    const int c = blockIdx.x, t = threadIdx.x, k = t >> 1;
    mp_unit_init(U);
    const int it0 = c * MP_CHUNK, it1 = it0 + MP_CHUNK < KZ_MP_ITERS ? it0 + MP_CHUNK : KZ_MP_ITERS;
    for (int it = it0; it < it1; ++it) {
        bool has_add;
        const int s0 = mp_step_of_iter(it, has_add);
        const int cnt = MP_NG * (has_add ? 2 : 1);
        const Fp12* src = part + (size_t)s0 * MP_NG;      // the addition step's partials follow the doubling step's
        if (t < 12) { if (t & 1) fit.c[k].c1 = src[0].c[k].c1; else fit.c[k].c0 = src[0].c[k].c0; }
        __syncthreads();
        for (int i = 1; i < cnt; ++i) {
            if (t < 12) { if (t & 1) y.c[k].c1 = src[i].c[k].c1; else y.c[k].c0 = src[i].c[k].c0; }
            __syncthreads();
            mp_mul(U, fit, fit, y);
        }
        if (it == it0) {
            if (t < 12) { if (t & 1) acc.c[k].c1 = fit.c[k].c1; else acc.c[k].c0 = fit.c[k].c0; }
            __syncthreads();
        } else {
            mp_mul(U, acc, acc, acc);
            mp_mul(U, acc, acc, fit);
        }
    }
    if (t < 12) { if (t & 1) G[c].c[k].c1 = acc.c[k].c1; else G[c].c[k].c0 = acc.c[k].c0; }
}
// The serial part: Miller accumulation f <- f^(2^s) G_c, conjugation (x < 0), inversion-free final check.
__global__ void __launch_bounds__(128) k_mp_check(const Fp12* __restrict__ G, int* __restrict__ result) {
    __shared__ MpScratch S;
    const int t = threadIdx.x, k = t >> 1;
    mp_unit_init(S.U);
    if (t < 12) { if (t & 1) S.f.c[k].c1 = G[0].c[k].c1; else S.f.c[k].c0 = G[0].c[k].c0; }
    __syncthreads();
    for (int c = 1; c < MP_NCHUNK; ++c) {
        const int nsq = (c + 1) * MP_CHUNK <= KZ_MP_ITERS ? MP_CHUNK : KZ_MP_ITERS - c * MP_CHUNK;
        Fp pre;                                                       // next factor: the load hides under the squarings
        if (t < 12) pre = (t & 1) ? G[c].c[k].c1 : G[c].c[k].c0;
        for (int q = 0; q < nsq; ++q) {
            mp_mul_products(S.U, t, S.f, S.f);
            __syncthreads();
            mp_mul_fold(S.U, t, S.f);
            if (q == 0 && t < 12) { if (t & 1) S.fb.c[k].c1 = pre; else S.fb.c[k].c0 = pre; }
            __syncthreads();
        }
        mp_mul_products(S.U, t, S.f, S.fb);
        __syncthreads();
        mp_mul_fold(S.U, t, S.f);
        __syncthreads();
    }
    coop_conj(S.f, S.f);
    mp_final_check(S);
    if (t == 0) *result = S.result;
}
void launch_mp_check(cudaStream_t s, const G2Lines* tab, const MpCoef* coef, Fp12* part, Fp12* F, int* result) {
    k_mp_lines<<<dim3(KZ_N_LINES, MP_NG), 128 * MP_UNITS, 0, s>>>(tab, coef, part);
    KZ_COUNT_LAUNCH();
    k_mp_merge<<<MP_NCHUNK, 128, 0, s>>>(part, F);
    KZ_COUNT_LAUNCH();
    k_mp_check<<<1, 128, 0, s>>>(F, result);
    KZ_COUNT_LAUNCH();
}
// A = sum_t 2^(4t) (sum of the shards' A-side terms), B likewise from the B-side terms (already negated): artefacts only
__global__ void __launch_bounds__(32) k_mp_ab(const G1Xyzz* __restrict__ terms_in, int n_shards, G1Jac* __restrict__ AB) {
    __shared__ Fp qsm[8 * KZ_QUAD_SLOTS];
    const int qi = threadIdx.x >> 2;
    if (qi >= 2) return;
    Quad q = quad_make(qsm, qi);
    G1Xyzz acc = xyzz_inf();
    for (int t = KZ_MP_TERMS - 1; t >= 0; --t) {
        for (int u = 0; u < KZ_MP_G; ++u) acc = quad_xyzz_dbl(q, acc);
        for (int g = 0; g < n_shards; ++g) acc = quad_xyzz_add(q, acc, terms_in[(size_t)g * KZ_MP_PAIRS + qi * KZ_MP_TERMS + t]);
    }
    if (q.ql == 0) AB[qi] = xyzz_to_jac(acc);
}
void launch_mp_ab(cudaStream_t s, const G1Xyzz* terms_in, int n_shards, G1Jac* AB) {
    k_mp_ab<<<1, 32, 0, s>>>(terms_in, n_shards, AB);
    KZ_COUNT_LAUNCH();
}
size_t mp_part_entries() { return (size_t)KZ_N_LINES * MP_NG; }

// ------------------------------------------------------------------ wire format of a shard's terms (kzgb200.h KZGB_TERMS_BYTES)
// 66 terms x (X | Y | ZZ | ZZZ), 48 B big-endian canonical each, then sum r_i y_i of the shard (32 B big-endian).
__global__ void __launch_bounds__(128) k_mp_terms_to_wire(const G1Xyzz* __restrict__ terms, const u32* __restrict__ sum_ry, u8* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * KZ_MP_PAIRS) fp_to_be(out + 48 * i, (&terms[0].X)[i]);
    else if (i == 4 * KZ_MP_PAIRS) {
        Fr v;
        for (int k = 0; k < 8; ++k) v.v[k] = sum_ry[k];
        fr_raw_to_be(out + 48 * 4 * KZ_MP_PAIRS, v);
    }
}
void launch_mp_terms_to_wire(cudaStream_t s, const G1Xyzz* terms, const uint32_t* sum_ry, uint8_t* out) {
    k_mp_terms_to_wire<<<(4 * KZ_MP_PAIRS + 1 + 127) / 128, 128, 0, s>>>(terms, sum_ry, out);
    KZ_COUNT_LAUNCH();
}
// n_shards wire records -> terms_in[g * 66 + p] (Montgomery), sum_ry_total = sum of the shards' sum r_i y_i;
// *bad += coordinates >= p or scalars >= r
__global__ void __launch_bounds__(128) k_mp_terms_from_wire(const u8* __restrict__ in, int n_shards, G1Xyzz* __restrict__ terms,
                                                             u32* __restrict__ sum_ry_total, u32* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, per = 4 * KZ_MP_PAIRS;
    if (i < per * n_shards) {
        const int g = i / per, k = i - g * per;
        Fp v;
        if (!fp_from_be(v, in + (size_t)g * (48 * per + 32) + 48 * k)) { atomicAdd(bad, 1u); v = fp_zero(); }
        (&terms[(size_t)g * KZ_MP_PAIRS].X)[k] = v;
    } else if (i == per * n_shards) {
        Fr acc = fr_zero();
        for (int g = 0; g < n_shards; ++g) {
            Fr v;
            fr_raw_from_be(v, in + (size_t)g * (48 * per + 32) + 48 * per);
            if (!fr_raw_is_canonical(v)) { atomicAdd(bad, 1u); continue; }
            acc = fr_add(acc, v);
        }
        for (int k = 0; k < 8; ++k) sum_ry_total[k] = acc.v[k];
    }
}
void launch_mp_terms_from_wire(cudaStream_t s, const uint8_t* in, int n_shards, G1Xyzz* terms, uint32_t* sum_ry_total, uint32_t* bad) {
    const int n = 4 * KZ_MP_PAIRS * n_shards + 1;
    k_mp_terms_from_wire<<<(n + 127) / 128, 128, 0, s>>>(in, n_shards, terms, sum_ry_total, bad);
    KZ_COUNT_LAUNCH();
}
