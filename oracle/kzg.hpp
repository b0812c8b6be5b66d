// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  KZG batch verification semantics on the CPU:
// decompress+validate, batched Fiat-Shamir, three MSMs, two-pairing check, synthetic generator.
// Semantics: BASELINE.json:5 and SURVEY.md App. B (upstream reference is LICENSE-only,
// /root/reference/LICENSE:1-201 -- nothing to follow).  Multithreaded with std::thread.
#pragma once
#include <array>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#include "curve.hpp"
#include "sha256.hpp"

namespace orc {

constexpr size_t CHUNK = 128;
constexpr u64 FIELD_ELEMENTS_PER_BLOB = 4096;

// ------------------------------------------------------------------ threads
inline void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t)>& fn, size_t grain = 1) {
    if (threads <= 1 || n <= grain) { fn(0, n); return; }
    std::atomic<size_t> next{0};
    size_t step = (n + (size_t)threads * 8 - 1) / ((size_t)threads * 8);
    if (step < grain) step = grain;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&] {
            for (;;) {
                size_t b = next.fetch_add(step);
                if (b >= n) break;
                fn(b, b + step < n ? b + step : n);
            }
        });
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------ Fr helpers
inline bool fr_from_be(Fr& out, const u8* b) { return Fr::from_bytes_be(out, b); }
inline void scalar_raw(const Fr& a, u64 raw[4]) { a.to_raw(raw); }

// ------------------------------------------------------------------ Fiat-Shamir (App. B.4)
static const char TAG_LEAF[] = "KZGB200/leaf_v1_";
static const char TAG_CHUNK[] = "KZGB200/chunk_v1";
static const char TAG_ROOT[] = "KZGB200/root_v1_";
static const char TAG_R[] = "KZGB200/r_v1____";

inline void fs_leaf(u8 out[32], const u8* C, const u8* z, const u8* y, const u8* pi) {
    Sha256 s;
    s.update(TAG_LEAF, 16); s.update(C, 48); s.update(z, 32); s.update(y, 32); s.update(pi, 48);
    s.final(out);
}
// digests for proofs [0,n): out has 32*ceil(n/CHUNK) bytes
inline void fs_chunk_digests(u8* out, const u8* C, const u8* z, const u8* y, const u8* pi, size_t n, int threads) {
    size_t nch = (n + CHUNK - 1) / CHUNK;
    parallel_for(nch, threads, [&](size_t b, size_t e) {
        for (size_t j = b; j < e; ++j) {
            Sha256 s;
            s.update(TAG_CHUNK, 16);
            size_t hi = (j + 1) * CHUNK < n ? (j + 1) * CHUNK : n;
            for (size_t i = j * CHUNK; i < hi; ++i) {
                u8 leaf[32];
                fs_leaf(leaf, C + 48 * i, z + 32 * i, y + 32 * i, pi + 48 * i);
                s.update(leaf, 32);
            }
            s.final(out + 32 * j);
        }
    });
}
inline void fs_root(u8 out[32], const u8* digests, size_t nch, u64 n_total) {
    Sha256 s;
    s.update(TAG_ROOT, 16);
    s.update_u64be(FIELD_ELEMENTS_PER_BLOB);
    s.update_u64be(n_total);
    s.update(digests, 32 * nch);
    s.final(out);
}
inline void fs_r(u8 out16[16], const u8 root[32], u64 i) {
    Sha256 s;
    s.update(TAG_R, 16); s.update(root, 32); s.update_u64be(i);
    u8 d[32];
    s.final(d);
    memcpy(out16, d, 16);
}

// ------------------------------------------------------------------ MSM (Pippenger, unsigned windows)
// Independent of the GPU's signed-digit scheme on purpose.  Scalars: raw little-endian u64[4].
inline int msm_window(size_t n) {
    int c = 1;
    while ((size_t(1) << (c + 2)) < n && c < 16) ++c;
    return c < 2 ? 2 : c;
}
inline G1J g1_madd(const G1J& a, const G1A& b) { return b.inf ? a : a.add(G1J::from_affine(b.x, b.y)); }

inline G1J msm(const G1A* pts, const u64 (*sc)[4], size_t n, int nbits, int threads) {
    if (n == 0) return G1J::inf();
    if (n < 8) {
        G1J acc = G1J::inf();
        for (size_t i = 0; i < n; ++i) acc = acc.add(pts[i].jac().mul(sc[i], 4));
        return acc;
    }
    int c = msm_window(n);
    int W = (nbits + c - 1) / c;
    std::vector<G1J> wsum(W);
    parallel_for(W, threads, [&](size_t wb, size_t we) {
        for (size_t w = wb; w < we; ++w) {
            std::vector<G1J> bucket((size_t(1) << c), G1J::inf());
            int lo = (int)w * c;
            for (size_t i = 0; i < n; ++i) {
                u64 d = sc[i][lo / 64] >> (lo % 64);
                if (lo % 64 + c > 64 && lo / 64 + 1 < 4) d |= sc[i][lo / 64 + 1] << (64 - lo % 64);
                d &= (u64(1) << c) - 1;
                if (d) bucket[d] = g1_madd(bucket[d], pts[i]);
            }
            G1J run = G1J::inf(), sum = G1J::inf();
            for (size_t k = (size_t(1) << c) - 1; k >= 1; --k) {
                run = run.add(bucket[k]);
                sum = sum.add(run);
            }
            wsum[w] = sum;
        }
    });
    G1J acc = G1J::inf();
    for (int w = W - 1; w >= 0; --w) {
        for (int k = 0; k < c; ++k) acc = acc.dbl();
        acc = acc.add(wsum[w]);
    }
    return acc;
}

// ------------------------------------------------------------------ batch verification
struct Artifacts {
    G1A S1, S2, S3, A, B;
    Fr sum_ry;
    u8 root[32];
    u64 n = 0;
    u32 n_bad_points = 0, n_bad_scalars = 0;
};
struct Setup {
    G1A g1;            // [tau^0]G1 as supplied
    G2A g2_0, g2_1;    // [tau^0]G2, [tau^1]G2
};

// Shard phase 1: validate + decompress + chunk digests.  Points kept for phase 2.
struct Shard {
    std::vector<G1A> c, pi;
    std::vector<Fr> z, y;
    size_t n = 0;
    u32 bad_points = 0, bad_scalars = 0;
};
inline void shard_phase1(Shard& sh, const u8* C, const u8* z, const u8* y, const u8* pi, size_t n, int threads, u8* digests) {
    sh.n = n;
    sh.c.resize(n); sh.pi.resize(n); sh.z.resize(n); sh.y.resize(n);
    std::atomic<u32> badp{0}, bads{0};
    parallel_for(n, threads, [&](size_t b, size_t e) {
        u32 lp = 0, ls = 0;
        for (size_t i = b; i < e; ++i) {
            lp += g1_decompress(sh.c[i], C + 48 * i) != ST_OK;
            lp += g1_decompress(sh.pi[i], pi + 48 * i) != ST_OK;
            ls += !fr_from_be(sh.z[i], z + 32 * i);
            ls += !fr_from_be(sh.y[i], y + 32 * i);
        }
        badp += lp; bads += ls;
    });
    sh.bad_points = badp; sh.bad_scalars = bads;
    fs_chunk_digests(digests, C, z, y, pi, n, threads);
}
// Shard phase 2: r_i, the three sums over this shard.  single => r_0 = 1 (verify_kzg_proof).
struct Partial {
    G1J s1, s2, s3;
    Fr sum_ry;
};
inline Partial shard_phase2(const Shard& sh, const u8 root[32], u64 offset, int threads, bool single = false) {
    size_t n = sh.n;
    std::vector<std::array<u64, 4>> r(n), rz(n);
    std::vector<Fr> ry(n);
    parallel_for(n, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            u8 be[32] = {0};
            if (single) be[31] = 1; else fs_r(be + 16, root, offset + i);
            Fr ri;
            fr_from_be(ri, be);          // 128-bit < r always
            ri.to_raw(r[i].data());
            (ri * sh.z[i]).to_raw(rz[i].data());
            ry[i] = ri * sh.y[i];
        }
    });
    Partial p;
    p.sum_ry = Fr::zero();
    for (size_t i = 0; i < n; ++i) p.sum_ry = p.sum_ry + ry[i];
    p.s1 = msm(sh.c.data(), (const u64(*)[4])r.data(), n, 128, threads);
    p.s2 = msm(sh.pi.data(), (const u64(*)[4])rz.data(), n, 255, threads);
    p.s3 = msm(sh.pi.data(), (const u64(*)[4])r.data(), n, 128, threads);
    return p;
}
inline bool combine_verify(Artifacts& art, const Setup& st, const Partial* parts, int np) {
    G1J s1 = G1J::inf(), s2 = G1J::inf(), s3 = G1J::inf();
    Fr sry = Fr::zero();
    for (int i = 0; i < np; ++i) {
        s1 = s1.add(parts[i].s1); s2 = s2.add(parts[i].s2); s3 = s3.add(parts[i].s3);
        sry = sry + parts[i].sum_ry;
    }
    u64 k[4];
    sry.to_raw(k);
    G1J a = s1.add(s2).add(st.g1.jac().mul(k, 4).neg());
    art.S1 = g1_affine(s1); art.S2 = g1_affine(s2); art.S3 = g1_affine(s3);
    art.A = g1_affine(a); art.B = g1_affine(s3.neg());
    art.sum_ry = sry;
    G1A P[2] = {art.A, art.B};
    G2A Q[2] = {st.g2_0, st.g2_1};
    return pairing_product_is_one(P, Q, 2);
}

// ------------------------------------------------------------------ synthetic inputs
inline void prng_block(u8 out[32], u64 seed, u64 stream, u64 k) {
    Sha256 s;
    s.update("kzgb200/prng", 12);
    s.update_u64be(seed); s.update_u64be(stream); s.update_u64be(k);
    s.final(out);
}
// 512-bit big-endian integer mod r
inline Fr fr_from_512(const u8 b[64]) {
    // value = hi*2^256 + lo ; both halves reduced by interpreting as raw then Montgomery-multiplying
    auto half = [](const u8* p) {
        u64 raw[4];
        for (int i = 0; i < 4; ++i) {
            u64 v = 0;
            for (int k = 0; k < 8; ++k) v = v << 8 | p[8 * (3 - i) + k];
            raw[i] = v;
        }
        // raw may be >= r: from_raw computes raw*R^2/R = raw*R mod r correctly for any raw < 2^256
        return Fr::from_raw(raw);
    };
    Fr two256 = Fr::from_raw(fr_params().one);     // canonical value of (2^256 mod r), lifted to Montgomery
    return half(b) * two256 + half(b + 32);
}
inline Fr prng_fr(u64 seed, u64 stream, u64 idx) {
    u8 b[64];
    prng_block(b, seed, stream, 2 * idx);
    prng_block(b + 32, seed, stream, 2 * idx + 1);
    return fr_from_512(b);
}
inline Fr test_tau() {
    Sha256 s;
    s.update("kzgb200/insecure-test-tau", 25);
    u8 d[32];
    s.final(d);
    u8 b[64] = {0};
    memcpy(b + 32, d, 32);
    return fr_from_512(b);
}
enum : u64 { STREAM_A = 1, STREAM_Z = 2, STREAM_Y = 3, STREAM_POLY = 7, STREAM_PLANT = 99 };

// fixed-base table: tab[w][j-1] = j * 2^(8w) * G, w<32, j in 1..255
struct FixedBase {
    std::vector<G1A> tab;
    explicit FixedBase(const G1A& g) {
        std::vector<G1J> j(32 * 255);
        G1J base = g.jac();
        for (int w = 0; w < 32; ++w) {
            G1J acc = base;
            for (int k = 0; k < 255; ++k) { j[w * 255 + k] = acc; acc = acc.add(base); }
            base = acc;   // 256 * base
        }
        tab.resize(j.size());
        g1_batch_affine(j.data(), tab.data(), j.size());
    }
    G1J mul(const Fr& k) const {
        u64 raw[4];
        k.to_raw(raw);
        G1J acc = G1J::inf();
        for (int w = 0; w < 32; ++w) {
            unsigned d = (raw[w / 8] >> (8 * (w % 8))) & 0xFF;
            if (d) acc = g1_madd(acc, tab[w * 255 + d - 1]);
        }
        return acc;
    }
};
inline const FixedBase& g1_fixed_base() {
    static const FixedBase fb(g1_generator());
    return fb;
}
inline void synth_instance(u64 seed, u64 offset, size_t n, u8* C, u8* z, u8* y, u8* pi, int threads) {
    const FixedBase& fb = g1_fixed_base();
    Fr tau = test_tau();
    parallel_for(n, threads, [&](size_t b, size_t e) {
        size_t m = e - b;
        std::vector<Fr> a(m), zz(m), yy(m), den(m), pre(m);
        Fr acc = Fr::one();
        for (size_t i = 0; i < m; ++i) {
            a[i] = prng_fr(seed, STREAM_A, offset + b + i);
            zz[i] = prng_fr(seed, STREAM_Z, offset + b + i);
            yy[i] = prng_fr(seed, STREAM_Y, offset + b + i);
            den[i] = tau - zz[i];
            pre[i] = acc;
            acc = acc * den[i];       // tau == z has probability 2^-255; inv(0)=0 then yields q=0
        }
        Fr inv = acc.inv();
        std::vector<G1J> pj(2 * m);
        for (size_t i = m; i-- > 0;) {
            Fr di = inv * pre[i];
            inv = inv * den[i];
            pj[2 * i] = fb.mul(a[i]);
            pj[2 * i + 1] = fb.mul((a[i] - yy[i]) * di);
        }
        std::vector<G1A> pa(2 * m);
        g1_batch_affine(pj.data(), pa.data(), 2 * m);
        for (size_t i = 0; i < m; ++i) {
            g1_compress(C + 48 * (b + i), pa[2 * i]);
            g1_compress(pi + 48 * (b + i), pa[2 * i + 1]);
            zz[i].to_bytes_be(z + 32 * (b + i));
            yy[i].to_bytes_be(y + 32 * (b + i));
        }
    }, 64);
}
// real-polynomial instances (config BJ:7): degree+1 random coefficients per proof
inline void synth_instance_poly(u64 seed, size_t n, size_t ncoef, u8* C, u8* z, u8* y, u8* pi, int threads) {
    const FixedBase& fb = g1_fixed_base();
    Fr tau = test_tau();
    parallel_for(n, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            Fr zi = prng_fr(seed, STREAM_Z, i);
            Fr ft = Fr::zero(), fz = Fr::zero();
            for (size_t j = ncoef; j-- > 0;) {
                Fr c = prng_fr(seed, STREAM_POLY, i * ncoef + j);
                ft = ft * tau + c;
                fz = fz * zi + c;
            }
            Fr q = (ft - fz) * (tau - zi).inv();
            g1_compress(C + 48 * i, g1_affine(fb.mul(ft)));
            g1_compress(pi + 48 * i, g1_affine(fb.mul(q)));
            zi.to_bytes_be(z + 32 * i);
            fz.to_bytes_be(y + 32 * i);
        }
    });
}
inline void synth_setup(u8* g1, size_t n1, u8* g2, size_t n2) {
    Fr tau = test_tau();
    Fr pw = Fr::one();
    for (size_t i = 0; i < n1 || i < n2; ++i) {
        u64 k[4];
        pw.to_raw(k);
        if (i < n1) g1_compress(g1 + 48 * i, g1_affine(g1_generator().jac().mul(k, 4)));
        if (i < n2) {
            G2J q = G2J::from_affine(g2_generator().x, g2_generator().y).mul(k, 4);
            G2A a{Fp2::zero(), Fp2::zero(), true};
            if (q.to_affine(a.x, a.y)) a.inf = false;
            g2_compress(g2 + 96 * i, a);
        }
        pw = pw * tau;
    }
}

}  // namespace orc
