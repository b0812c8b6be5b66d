// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PeerDAS-shaped cell batch (BASELINE.json config[4]):
// multi-point KZG openings on cosets of size 64, universal verification equation.  SPEC: DESIGN.md "Cell batch".
// Upstream reference is LICENSE-only (/root/reference/LICENSE:1-201); semantics pinned by oracle/pymodel.
#pragma once
#include <mutex>

#include "kzg.hpp"

namespace orc {

constexpr size_t N_EXT = 8192, N_CELLS = 128, CELL_LEN = 64, CELL_BYTES = 2048;
constexpr u64 STREAM_BLOB = 11;

inline Fr fr_pow_u64(Fr a, u64 e) {
    u64 limbs[1] = {e};
    return a.pow(limbs, 1);
}
// omega = 7^((r-1)/8192): primitive 8192-th root of unity
inline const Fr& omega_ext() {
    static const Fr w = [] {
        u64 e[4], onev[4] = {1};
        sub_raw<4>(e, fr_params().mod, onev);
        for (int i = 0; i < 4; ++i) e[i] = (e[i] >> 13) | (i < 3 ? e[i + 1] << 51 : 0);
        return Fr::from_u64(7).pow(e, 4);
    }();
    return w;
}
inline unsigned brp(unsigned v, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}
inline Fr coset_shift(unsigned c) { return fr_pow_u64(omega_ext(), brp(c, 7)); }

// in-place radix-2 NTT of size n (power of two) with root w of order n: a[k] <- sum_j a[j] w^(jk)
inline void ntt(std::vector<Fr>& a, const Fr& w) {
    size_t n = a.size();
    int lg = 0;
    while ((size_t(1) << lg) < n) ++lg;
    for (size_t i = 0; i < n; ++i) {
        size_t j = brp((unsigned)i, lg);
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        Fr wl = fr_pow_u64(w, n / len);
        for (size_t i = 0; i < n; i += len) {
            Fr t = Fr::one();
            for (size_t k = 0; k < len / 2; ++k) {
                Fr u = a[i + k], v = a[i + k + len / 2] * t;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
                t = t * wl;
            }
        }
    }
}
// coefficients a_0..a_63 of the interpolation polynomial of cell c: I(h_c w64^j) = ys[j]
inline std::vector<Fr> interp_coeffs(unsigned c, const Fr* ys) {
    std::vector<Fr> a(ys, ys + CELL_LEN);
    Fr w64 = fr_pow_u64(omega_ext(), N_EXT / CELL_LEN);
    ntt(a, w64.inv());
    Fr hinv = coset_shift(c).inv(), scale = Fr::from_u64(CELL_LEN).inv();
    for (size_t i = 0; i < CELL_LEN; ++i) {
        a[i] = a[i] * scale;
        scale = scale * hinv;
    }
    return a;
}

static const char TAG_CELL[] = "KZGB200/cell_v1_";
static const char TAG_COMM[] = "KZGB200/comm_v1_";
static const char TAG_CROOT[] = "KZGB200/croot_v1";

inline void cell_root(u8 root[32], const u8* comms, size_t nc, const u32* ci, const u32* xi, const u8* cells, const u8* proofs,
                      size_t m, int threads) {
    std::vector<u8> leaves(32 * m);
    parallel_for(m, threads, [&](size_t b, size_t e) {
        for (size_t k = b; k < e; ++k) {
            Sha256 s;
            s.update(TAG_CELL, 16);
            s.update_u64be(ci[k]); s.update_u64be(xi[k]);
            s.update(cells + CELL_BYTES * k, CELL_BYTES);
            s.update(proofs + 48 * k, 48);
            s.final(&leaves[32 * k]);
        }
    });
    size_t nch = (m + CHUNK - 1) / CHUNK;
    std::vector<u8> dig(32 * nch);
    for (size_t j = 0; j < nch; ++j) {
        Sha256 s;
        s.update(TAG_CHUNK, 16);
        size_t hi = (j + 1) * CHUNK < m ? (j + 1) * CHUNK : m;
        s.update(&leaves[32 * j * CHUNK], 32 * (hi - j * CHUNK));
        s.final(&dig[32 * j]);
    }
    u8 cdig[32];
    Sha256 sc;
    sc.update(TAG_COMM, 16);
    sc.update(comms, 48 * nc);
    sc.final(cdig);
    Sha256 s;
    s.update(TAG_CROOT, 16);
    s.update_u64be(nc); s.update_u64be(m);
    s.update(cdig, 32);
    s.update(dig.data(), dig.size());
    s.final(root);
}

struct CellSetup {
    std::vector<G1A> g1;      // [tau^j]G1, j < 64
    G2A g2_0, g2_64;
    bool ready = false;
};

// returns KZGB_OK(0) / KZGB_BADARGS(1); verdict in ok; artefacts A, B, root
inline int verify_cells(bool& ok, Artifacts& art, const CellSetup& st, const u8* comms, size_t nc, const u32* ci, const u32* xi,
                        const u8* cells, const u8* proofs, size_t m, int threads) {
    ok = false;
    if (!st.ready || m == 0 || nc == 0) return 1;
    std::vector<G1A> C(nc), PI(m);
    std::vector<Fr> ys(m * CELL_LEN);
    std::atomic<u32> bad{0};
    parallel_for(nc, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) bad += g1_decompress(C[i], comms + 48 * i) != ST_OK;
    });
    parallel_for(m, threads, [&](size_t b, size_t e) {
        u32 lb = 0;
        for (size_t k = b; k < e; ++k) {
            lb += g1_decompress(PI[k], proofs + 48 * k) != ST_OK;
            lb += xi[k] >= N_CELLS;
            lb += ci[k] >= nc;
            for (size_t j = 0; j < CELL_LEN; ++j) lb += !Fr::from_bytes_be(ys[k * CELL_LEN + j], cells + CELL_BYTES * k + 32 * j);
        }
        bad += lb;
    });
    art = Artifacts();
    art.n = m;
    art.n_bad_points = bad;
    if (bad) return 1;
    cell_root(art.root, comms, nc, ci, xi, cells, proofs, m, threads);
    // scalars
    std::vector<std::array<u64, 4>> r(m), rh(m);
    std::vector<Fr> rfr(m);
    std::vector<std::vector<Fr>> partial_s((size_t)threads > 0 ? threads : 1, std::vector<Fr>(CELL_LEN, Fr::zero()));
    std::atomic<int> slot{0};
    parallel_for(m, threads, [&](size_t b, size_t e) {
        std::vector<Fr> acc(CELL_LEN, Fr::zero());
        for (size_t k = b; k < e; ++k) {
            u8 be[32] = {0};
            fs_r(be + 16, art.root, k);
            Fr rk;
            fr_from_be(rk, be);
            rfr[k] = rk;
            rk.to_raw(r[k].data());
            Fr h64 = fr_pow_u64(coset_shift(xi[k]), CELL_LEN);
            (rk * h64).to_raw(rh[k].data());
            std::vector<Fr> a = interp_coeffs(xi[k], &ys[k * CELL_LEN]);
            for (size_t i = 0; i < CELL_LEN; ++i) acc[i] = acc[i] + rk * a[i];
        }
        int s = slot.fetch_add(1) % (int)partial_s.size();
        static std::mutex mu;
        std::lock_guard<std::mutex> g(mu);
        for (size_t i = 0; i < CELL_LEN; ++i) partial_s[s][i] = partial_s[s][i] + acc[i];
    }, 16);
    std::vector<Fr> S(CELL_LEN, Fr::zero());
    for (auto& v : partial_s) for (size_t i = 0; i < CELL_LEN; ++i) S[i] = S[i] + v[i];
    std::vector<Fr> w(nc, Fr::zero());
    for (size_t k = 0; k < m; ++k) w[ci[k]] = w[ci[k]] + rfr[k];
    // A = sum w_i C_i - sum S_j T_j + sum (r_k h_k^64) pi_k ;  B = -sum r_k pi_k
    std::vector<std::array<u64, 4>> wraw(nc), sraw(CELL_LEN);
    for (size_t i = 0; i < nc; ++i) w[i].to_raw(wraw[i].data());
    for (size_t i = 0; i < CELL_LEN; ++i) (-S[i]).to_raw(sraw[i].data());
    G1J a = msm(C.data(), (const u64(*)[4])wraw.data(), nc, 255, threads);
    a = a.add(msm(st.g1.data(), (const u64(*)[4])sraw.data(), CELL_LEN, 255, threads));
    a = a.add(msm(PI.data(), (const u64(*)[4])rh.data(), m, 255, threads));
    G1J bsum = msm(PI.data(), (const u64(*)[4])r.data(), m, 128, threads);
    art.S1 = art.S2 = art.S3 = G1A::infinity();
    art.A = g1_affine(a);
    art.B = g1_affine(bsum.neg());
    art.sum_ry = Fr::zero();
    G1A P[2] = {art.A, art.B};
    G2A Q[2] = {st.g2_0, st.g2_64};
    ok = pairing_product_is_one(P, Q, 2);
    return 0;
}

// synthetic cells: n_blobs random polynomials with ncoef (<= 4096) coefficients, cells 0..cells_per_blob-1 of each
inline void synth_cells(u64 seed, size_t n_blobs, size_t cells_per_blob, size_t ncoef, u8* comms, u32* ci, u32* xi, u8* cells,
                        u8* proofs, int threads) {
    const FixedBase& fb = g1_fixed_base();
    Fr tau = test_tau();
    Fr tau64 = fr_pow_u64(tau, CELL_LEN);
    parallel_for(n_blobs, threads, [&](size_t b0, size_t b1) {
        for (size_t bi = b0; bi < b1; ++bi) {
            std::vector<Fr> ev(N_EXT, Fr::zero());
            Fr ft = Fr::zero();
            for (size_t j = ncoef; j-- > 0;) {
                ev[j] = prng_fr(seed, STREAM_BLOB, bi * 4096 + j);
                ft = ft * tau + ev[j];
            }
            ntt(ev, omega_ext());                       // ev[t] = f(omega^t)
            g1_compress(comms + 48 * bi, g1_affine(fb.mul(ft)));
            for (size_t c = 0; c < cells_per_blob; ++c) {
                size_t k = bi * cells_per_blob + c;
                Fr ys[CELL_LEN];
                for (size_t j = 0; j < CELL_LEN; ++j) {
                    ys[j] = ev[(brp((unsigned)c, 7) + 128 * j) % N_EXT];
                    ys[j].to_bytes_be(cells + CELL_BYTES * k + 32 * j);
                }
                std::vector<Fr> a = interp_coeffs((unsigned)c, ys);
                Fr it = Fr::zero();
                for (size_t i = CELL_LEN; i-- > 0;) it = it * tau + a[i];
                Fr den = tau64 - fr_pow_u64(coset_shift((unsigned)c), CELL_LEN);
                g1_compress(proofs + 48 * k, g1_affine(fb.mul((ft - it) * den.inv())));
                ci[k] = (u32)bi;
                xi[k] = (u32)c;
            }
        }
    });
}

}  // namespace orc
