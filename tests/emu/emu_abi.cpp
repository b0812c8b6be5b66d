// TEST INFRASTRUCTURE ONLY -- host emulation of the CUDA kernels' per-thread bodies.
// The device math headers of kzg_batch_verification_scheme_b200/csrc are compiled here by g++ with
// -DKZGB_EMU (portable limb loops instead of the inline-PTX blocks, `for` loops instead of thread
// indices) so the no-GPU CI can exercise the SAME decompression / subgroup / SHA / MSM / pairing logic
// that runs on the B200 and diff it against the oracle.  Nothing in the product library links this.
#include <algorithm>
#include <cfenv>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/kzgb200.h"
#include "../../kzg_batch_verification_scheme_b200/csrc/msm.cuh"
#include "../../kzg_batch_verification_scheme_b200/csrc/pairing.cuh"
#include "../../kzg_batch_verification_scheme_b200/csrc/sha256.cuh"
#include "../../kzg_batch_verification_scheme_b200/csrc/cells.cuh"
#include "../../tools/microbench/fpd.cuh"     // recorded negative result (FP64-limb multiplier): emulation test only, not in the product
#include "../../kzg_batch_verification_scheme_b200/csrc/blob.cuh"
#include "../../kzg_batch_verification_scheme_b200/csrc/mpair.cuh"
#include "../../kzg_batch_verification_scheme_b200/csrc/eip4844.cuh"

struct kzgb_ctx {
    G2Lines lines[2];
    G2Lines lines_cell[2];
    std::vector<G1Aff> cell_g1;
    std::vector<Fr> W;
    bool cell_ready = false;
    G1Jac ab[2];
    bool have_ab = false;
    G1Aff g1;
    kzgb_artifacts art;
    G1Jac sums[3];
    Fr sum_ry;
    bool have_sums = false;
    size_t sg_min = 0;             // batched subgroup check for batches of at least this many proofs (0 = never)
    std::vector<G2Lines> mp_tab;   // Horner-free pairing check: lines of [2^(4t)]G2, [2^(4t)][tau]G2, t < 33
};

// what the device keeps of one sum for the Horner-free pairing check: plan, bucket table, slice sums
struct EmuSum {
    MsmPlan plan;
    std::vector<G1Xyzz> buckets, slices;
    MpSumDesc desc() const { return {slices.data(), buckets.data(), plan.c, plan.W, plan.nbits}; }
};

static void words_from_be(u32* w, const u8* in, int nw) {
    for (int i = 0; i < nw; ++i) w[i] = (u32)in[4 * i] << 24 | in[4 * i + 1] << 16 | in[4 * i + 2] << 8 | in[4 * i + 3];
}
static void words_to_be(u8* out, const u32* w, int nw) {
    for (int i = 0; i < nw; ++i) { out[4 * i] = w[i] >> 24; out[4 * i + 1] = w[i] >> 16; out[4 * i + 2] = w[i] >> 8; out[4 * i + 3] = w[i]; }
}

// emulated MSM pipeline: digits -> sort -> bounds -> accumulate -> segments -> window sums -> combine
// sg_fail != null: also run the batched subgroup check on this sum's buckets (count of slice sums outside G1)
static G1Jac emu_msm(const Fp* pts, const u32* scalars, int nl, size_t m, int nbits, u32* sg_fail = nullptr, EmuSum* keep = nullptr) {
    if (nbits == 255) {      // GLV split exactly as the device pipeline does it
        std::vector<u32> zs(4 * 2 * m);
        std::vector<Fp> p2(2 * 2 * m);
        for (size_t i = 0; i < m; ++i) {
            u32 sc[8];
            for (int k = 0; k < 8; ++k) sc[k] = k < nl ? scalars[nl * i + k] : 0;
            glv_split(sc, &zs[4 * i], &zs[4 * (m + i)]);
            G1Aff p = load_point(pts, i), q = g1_endo(p);
            p2[2 * i] = p.x; p2[2 * i + 1] = p.y;
            p2[2 * (m + i)] = q.x; p2[2 * (m + i) + 1] = q.y;
        }
        return emu_msm(p2.data(), zs.data(), 4, 2 * m, 128, nullptr, keep);
    }
    MsmPlan plan = msm_make_plan(m, nbits);
    size_t N = m * (size_t)plan.W;
    std::vector<u32> keys(N), vals(N);
    for (size_t i = 0; i < m; ++i) {
        u32 sc[8];
        for (int k = 0; k < 8; ++k) sc[k] = k < nl ? scalars[nl * i + k] : 0;
        msm_digits_body(keys.data(), vals.data(), sc, i, m, plan);
    }
    std::vector<size_t> order(N);
    for (size_t i = 0; i < N; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    std::vector<u32> sk(N), sv(N);
    for (size_t i = 0; i < N; ++i) { sk[i] = keys[order[i]]; sv[i] = vals[order[i]]; }
    std::vector<u32> start(plan.total_buckets + 2);
    for (u32 b = 0; b <= plan.total_buckets + 1; ++b) start[b] = (u32)(std::lower_bound(sk.begin(), sk.end(), b) - sk.begin());
    std::vector<G1Xyzz> buckets(plan.total_buckets), wins(plan.W);
    // balanced two-pass accumulation with a short chunk so that buckets straddle many chunks
    {
        u32 L = 3, n_valid = start[plan.total_buckets], T = (u32)((N + L - 1) / L);
        for (auto& bk : buckets) bk = xyzz_inf();
        std::vector<G1Xyzz> head(T + 1), tail(T + 1);
        std::vector<u32> hk(T + 1), tk(T + 1), hf(T + 1);
        ChunkRecs R{head.data(), tail.data(), hk.data(), tk.data(), hf.data()};
        for (u32 t = 0; t < T; ++t) msm_chunk_pass1(pts, sk.data(), sv.data(), n_valid, L, t, buckets.data(), R);
        for (u32 t = 0; t < T; ++t) msm_chunk_pass2(T, t, buckets.data(), R);
        // cross-check against the plain one-thread-per-bucket body
        for (u32 b = 0; b < plan.total_buckets; ++b) {
            G1Xyzz ref = msm_bucket_body(pts, sv.data(), start[b], start[b + 1]);
            G1Jac a = xyzz_to_jac(ref), c2 = xyzz_to_jac(buckets[b]);
            bool same = jac_is_inf(a) ? jac_is_inf(c2) : (!jac_is_inf(c2) && aff_is_inf(jac_to_aff(jac_add(a, jac_neg(c2)))));
            if (!same) { fprintf(stderr, "emu: bucket %u differs\n", b); abort(); }
        }
    }
    // bucket reduction exactly as the device does it: run sums, row / column totals, slice sums, window Horner
    const SgLayout L = sg_layout(plan);
    std::vector<G1Xyzz> part((size_t)plan.W * L.stride), tot((size_t)plan.W * L.tstride), slices(plan.nbits);
    for (int w = 0; w < plan.W; ++w) {
        const SgWin g = sg_win(plan, w);
        for (u32 t = 0; t < g.jobs; ++t) part[(size_t)w * L.stride + t] = sg_run_sum(buckets.data() + plan.bucket_off[w], g, t);
        for (u32 t = 0; t < g.rows + g.cols; ++t) tot[(size_t)w * L.tstride + t] = sg_total(part.data() + (size_t)w * L.stride, g, t);
    }
    for (int sid = 0; sid < plan.nbits; ++sid) {
        const SgSlice sl = sg_slice(plan, sid);
        const SgWin g = sg_win(plan, sl.w);
        G1Xyzz acc = xyzz_inf();
        for (u32 lane = 0; lane < 5; ++lane)          // any lane count gives the same sum
            acc = xyzz_add(acc, sg_slice_part(tot.data() + (size_t)sl.w * L.tstride, buckets.data() + plan.bucket_off[sl.w], g, sl, lane, 5));
        slices[sid] = acc;
        if (sg_fail) {
            // the device runs the |x|^2 chains on quads (quad.cuh): both forms of the test must agree
            Quad q = quad_make(nullptr, 0);
            const bool in1 = sg_sum_in_g1(acc), in2 = quad_sum_in_g1(q, acc);
            if (in1 != in2) { fprintf(stderr, "emu: quad and one-thread subgroup tests differ on slice %d\n", sid); abort(); }
            if (!in1) *sg_fail += 1;
        }
    }
    for (int w = 0; w < plan.W; ++w) {
        const SgWin g = sg_win(plan, w);
        wins[w] = msm_window_from_slices(slices.data() + (size_t)w * plan.c, g.k, buckets[plan.bucket_off[w] + (1u << g.k) - 1u]);
        // cross-check against the definition sum_m m * B_m
        G1Xyzz run = xyzz_inf(), ref = xyzz_inf();
        for (u32 m = plan.nb[w]; m >= 1; --m) { run = xyzz_add(run, buckets[plan.bucket_off[w] + m - 1]); ref = xyzz_add(ref, run); }
        G1Jac a = xyzz_to_jac(ref), c2 = xyzz_to_jac(wins[w]);
        bool same = jac_is_inf(a) ? jac_is_inf(c2) : (!jac_is_inf(c2) && aff_is_inf(jac_to_aff(jac_add(a, jac_neg(c2)))));
        if (!same) { fprintf(stderr, "emu: window %d total differs\n", w); abort(); }
    }
    if (keep) { keep->plan = plan; keep->buckets = buckets; keep->slices = slices; }
    return msm_combine_body(wins.data(), plan.W, plan.c);
}

// ---- Horner-free pairing check (mpair.cuh) restated sequentially: the same terms, coefficients, line values and final
// check as k_mp_terms / k_mp_coefs / k_mp_lines / k_mp_check, without the chunked merging of Miller iterations (which
// only regroups the same product).  tab = 66 line tables, sums d1 + d2 on the A side, d3 (negated) on the B side.
static bool emu_mp_setup(std::vector<G2Lines>& tab, const u8* g2_two_points) {
    tab.resize(KZ_MP_PAIRS);
    for (int b = 0; b < 2; ++b) {
        G2Aff q;
        if (!g2_decompress(q, g2_two_points + 96 * b)) return false;
        Fp2 la, lb;
        for (int t = 0; t < KZ_MP_TERMS; ++t) {
            G2Aff out;
            if (!g2_mul_xabs(out, q, &tab[b * KZ_MP_TERMS + t])) return false;
            for (int u = 0; u < KZ_MP_G && t + 1 < KZ_MP_TERMS; ++u)
                if (!g2_dbl_step(q, la, lb)) return false;
        }
    }
    return true;
}
static bool emu_mp_check(const std::vector<G2Lines>& tab, const MpSumDesc& d1, const MpSumDesc& d2, const MpSumDesc& d3) {
    Quad q = quad_make(nullptr, 0);
    std::vector<MpCoef> coef(KZ_MP_PAIRS);
    for (int t = 0; t < KZ_MP_TERMS; ++t) {
        G1Xyzz a = quad_xyzz_add(q, mp_term(q, d1, t), mp_term(q, d2, t));
        coef[t] = mp_coef_of(q, a);
        coef[KZ_MP_TERMS + t] = mp_coef_of(q, xyzz_neg(mp_term(q, d3, t)));
    }
    static MpScratch S;
    mp_unit_init(S.U);
    auto line_product = [&](int s) {
        Fp12 acc;
        for (int p = 0; p < KZ_MP_PAIRS; ++p) {
            Fp12 l;
            for (int ci = 0; ci < 12; ++ci) {
                Fp v = mp_line_coeff(tab[p], s, coef[p], ci);
                if (ci & 1) l.c[ci >> 1].c1 = v; else l.c[ci >> 1].c0 = v;
            }
            if (p == 0) acc = l; else mp_mul(S.U, acc, acc, l);
        }
        return acc;
    };
    Fp12 f;
    for (int it = 0; it < KZ_MP_ITERS; ++it) {
        bool has_add;
        const int s0 = mp_step_of_iter(it, has_add);
        Fp12 F = line_product(s0);
        if (has_add) { Fp12 G = line_product(s0 + 1); mp_mul(S.U, F, F, G); }
        if (it == 0) f = F;
        else { mp_mul(S.U, f, f, f); mp_mul(S.U, f, f, F); }
    }
    coop_conj(S.f, f);
    mp_final_check(S);
    return S.result == 1;
}

extern "C" {

const char* kzgb_version(void) { return "kzgb200-emu (tests only)"; }

kzgb_ret kzgb_ctx_create(kzgb_ctx** out, const uint8_t* g1m, size_t n1, const uint8_t* g2m, size_t n2, const int*, int, size_t) {
    if (!out || !g1m || !g2m || n1 < 1 || n2 < 2) return KZGB_BADARGS;
    kzgb_ctx* c = new kzgb_ctx();
    u32 w[12];
    words_from_be(w, g1m, 12);
    bool ok = g1_decompress_validate(c->g1, w) == ST_OK && !aff_is_inf(c->g1);
    ok = ok && g2_setup_point(c->lines[0], g2m) && g2_setup_point(c->lines[1], g2m + 96);
    ok = ok && emu_mp_setup(c->mp_tab, g2m);
    if (!ok) { delete c; return KZGB_BADARGS; }
    if (n1 >= 64 && n2 >= 65) {
        c->cell_g1.resize(64);
        for (int j = 0; j < 64 && ok; ++j) {
            words_from_be(w, g1m + 48 * j, 12);
            ok = g1_decompress_validate(c->cell_g1[j], w) == ST_OK;
        }
        c->lines_cell[0] = c->lines[0];
        ok = ok && g2_setup_point(c->lines_cell[1], g2m + 96 * 64);
        if (!ok) { delete c; return KZGB_BADARGS; }
        c->W.resize(KZ_N_EXT);
        Fr acc = fr_const(FR_ONE);
        for (u32 t = 0; t < KZ_N_EXT; ++t) { c->W[t] = acc; acc = fr_mul(acc, fr_const(FR_OMEGA_INV)); }
        c->cell_ready = true;
    }
    *out = c;
    return KZGB_OK;
}
void kzgb_ctx_free(kzgb_ctx* c) { delete c; }

kzgb_ret kzgb_g1_decompress_batch(uint8_t* aff, uint8_t* st, const uint8_t* in, size_t m, kzgb_ctx*) {
    for (size_t i = 0; i < m; ++i) {
        u32 w[12];
        words_from_be(w, in + 48 * i, 12);
        G1Aff p;
        st[i] = (u8)g1_decompress_validate(p, w);
        aff_to_be96(aff + 96 * i, p);
    }
    return KZGB_OK;
}

static void emu_digests(std::vector<u8>& dig, const u8* C, const u8* z, const u8* y, const u8* pi, size_t n) {
    std::vector<u32> leaves(8 * n);
    for (size_t i = 0; i < n; ++i) {
        u32 cw[12], pw[12], zw[8], yw[8];
        words_from_be(cw, C + 48 * i, 12); words_from_be(pw, pi + 48 * i, 12);
        words_from_be(zw, z + 32 * i, 8); words_from_be(yw, y + 32 * i, 8);
        fs_leaf_words(&leaves[8 * i], cw, zw, yw, pw);
    }
    size_t nch = (n + KZ_FS_CHUNK - 1) / KZ_FS_CHUNK;
    dig.resize(32 * nch);
    for (size_t j = 0; j < nch; ++j) {
        u32 h[8];
        size_t lo = j * KZ_FS_CHUNK;
        fs_chunk_words(h, &leaves[8 * lo], (u32)std::min<size_t>(KZ_FS_CHUNK, n - lo));
        words_to_be(&dig[32 * j], h, 8);
    }
}
// root hash through the device-side compression function (message = tag | deg | n | digests)
static void emu_root(u8 root[32], const std::vector<u8>& dig, u64 n) {
    std::vector<u8> msg;
    const char* tag = "KZGB200/root_v1_";
    msg.insert(msg.end(), tag, tag + 16);
    for (int i = 0; i < 8; ++i) msg.push_back((u8)((u64)4096 >> (56 - 8 * i)));
    for (int i = 0; i < 8; ++i) msg.push_back((u8)(n >> (56 - 8 * i)));
    msg.insert(msg.end(), dig.begin(), dig.end());
    u64 bits = msg.size() * 8;
    msg.push_back(0x80);
    while (msg.size() % 64 != 56) msg.push_back(0);
    for (int i = 0; i < 8; ++i) msg.push_back((u8)(bits >> (56 - 8 * i)));
    u32 h[8];
    sha256_init(h);
    for (size_t b = 0; b < msg.size(); b += 64) {
        u32 w[16];
        words_from_be(w, &msg[b], 16);
        sha256_compress(h, w);
    }
    words_to_be(root, h, 8);
}

kzgb_ret kzgb_fs_challenges(uint8_t root[32], uint8_t* r_out, const uint8_t* C, const uint8_t* z, const uint8_t* y,
                            const uint8_t* pi, size_t n, kzgb_ctx*) {
    std::vector<u8> dig;
    emu_digests(dig, C, z, y, pi, n);
    emu_root(root, dig, n);
    u32 rw[8];
    words_from_be(rw, root, 8);
    for (size_t i = 0; i < n; ++i) {
        u32 r[4];
        fs_r_limbs(r, rw, i);
        u32 be[4] = {r[3], r[2], r[1], r[0]};
        words_to_be(r_out + 16 * i, be, 4);
    }
    return KZGB_OK;
}

static kzgb_ret emu_verify(bool* ok, const u8* C, const u8* z, const u8* y, const u8* pi, size_t n, kzgb_ctx* c, bool single) {
    *ok = false;
    if (!n) return KZGB_BADARGS;
    memset(&c->art, 0, sizeof c->art);
    c->art.n = n;
    c->have_sums = false;
    const bool sg_batch = !single && c->sg_min && n >= c->sg_min && n >= 2;
    std::vector<Fp> pts(2 * (2 * n + 1));
    u32 badp = 0, bads = 0, sg_fail = 0;
    for (size_t i = 0; i < 2 * n; ++i) {
        u32 w[12];
        words_from_be(w, i < n ? C + 48 * i : pi + 48 * (i - n), 12);
        G1Aff p;
        badp += g1_decompress_validate(p, w, !sg_batch) != ST_OK;
        pts[2 * i] = p.x; pts[2 * i + 1] = p.y;
    }
    std::vector<u8> dig;
    emu_digests(dig, C, z, y, pi, n);
    u8 root[32] = {0};
    if (!single) emu_root(root, dig, n);
    memcpy(c->art.root, root, 32);
    u32 rw[8];
    words_from_be(rw, root, 8);
    std::vector<u32> r(4 * n), rz(8 * (n + 1));
    Fr sum = fr_zero();
    for (size_t i = 0; i < n; ++i) {
        Fr ri = fr_zero(), zr, yr;
        if (single) ri.v[0] = 1; else fs_r_limbs(ri.v, rw, i);
        fr_raw_from_be(zr, z + 32 * i); fr_raw_from_be(yr, y + 32 * i);
        bads += !fr_raw_is_canonical(zr); bads += !fr_raw_is_canonical(yr);
        Fr v = fr_mul(ri, fr_to_mont(zr));
        sum = fr_add(sum, fr_mul(ri, fr_to_mont(yr)));
        for (int k = 0; k < 4; ++k) r[4 * i + k] = ri.v[k];
        for (int k = 0; k < 8; ++k) rz[8 * i + k] = v.v[k];
    }
    if (sg_batch && (badp || bads)) {
        // the device still runs the sums and their slice check on the well-formed points; an off-subgroup point
        // among them fails a slice and the per-point fallback counts it
        for (size_t i = 0; i < 2 * n; ++i) {
            G1Aff p = load_point(pts.data(), i);
            badp += !aff_is_inf(p) && !g1_in_subgroup(p);
        }
    }
    c->art.n_bad_points = badp; c->art.n_bad_scalars = bads;
    if (badp || bads) return KZGB_BADARGS;
    Fr neg = fr_neg(sum);
    for (int k = 0; k < 8; ++k) rz[8 * n + k] = neg.v[k];
    pts[2 * 2 * n] = c->g1.x; pts[2 * 2 * n + 1] = c->g1.y;
    EmuSum k1, k2, k3;
    c->sums[0] = emu_msm(pts.data(), r.data(), 4, n, 128, sg_batch ? &sg_fail : nullptr, &k1);
    c->sums[2] = emu_msm(pts.data() + 2 * n, r.data(), 4, n, 128, sg_batch ? &sg_fail : nullptr, &k3);
    if (sg_fail) {           // a slice sum left G1: the per-point check names the offenders
        for (size_t i = 0; i < 2 * n; ++i) {
            G1Aff p = load_point(pts.data(), i);
            badp += !aff_is_inf(p) && !g1_in_subgroup(p);
        }
        c->art.n_bad_points = badp;
        return KZGB_BADARGS;
    }
    c->sums[1] = emu_msm(pts.data() + 2 * n, rz.data(), 8, n + 1, 255, nullptr, &k2);
    c->sum_ry = sum;
    c->have_sums = true;
    G1Jac AB[2] = {jac_add(c->sums[0], c->sums[1]), jac_neg(c->sums[2])};
    PairScratch S;
    coop_pairing_check(S, c->lines, AB);
    *ok = S.result == 1;
    // the product's default path: multi-pairing over the slice-sum terms, inversion-free final check -- same verdict
    const bool ok2 = emu_mp_check(c->mp_tab, k1.desc(), k2.desc(), k3.desc());
    if (ok2 != *ok) { fprintf(stderr, "emu: Horner-free pairing check (%d) and two-pairing kernel (%d) disagree\n", (int)ok2, (int)*ok); abort(); }
    return KZGB_OK;
}
kzgb_ret verify_kzg_proof(bool* ok, const uint8_t C[48], const uint8_t z[32], const uint8_t y[32], const uint8_t pi[48], kzgb_ctx* c) {
    return emu_verify(ok, C, z, y, pi, 1, c, true);
}
kzgb_ret verify_kzg_proof_batch(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n, kzgb_ctx* c) {
    return emu_verify(ok, C, z, y, pi, n, c, false);
}
kzgb_ret kzgb_last_artifacts(kzgb_ctx* c, kzgb_artifacts* out) {
    if (c->have_ab) {
        aff_to_be96(c->art.A, jac_to_aff(c->ab[0]));
        aff_to_be96(c->art.B, jac_to_aff(c->ab[1]));
    }
    if (c->have_sums) {
        u32 k[8];
        for (int i = 0; i < 8; ++i) k[i] = c->sum_ry.v[i];
        G1Jac s2 = jac_add(c->sums[1], jac_mul_limbs(jac_from_aff(c->g1), k, 8));
        aff_to_be96(c->art.S1, jac_to_aff(c->sums[0]));
        aff_to_be96(c->art.S2, jac_to_aff(s2));
        aff_to_be96(c->art.S3, jac_to_aff(c->sums[2]));
        aff_to_be96(c->art.A, jac_to_aff(jac_add(c->sums[0], c->sums[1])));
        aff_to_be96(c->art.B, jac_to_aff(jac_neg(c->sums[2])));
        fr_raw_to_be(c->art.sum_ry, c->sum_ry);
    }
    *out = c->art;
    return KZGB_OK;
}
kzgb_ret kzgb_g1_msm(uint8_t out[96], const uint8_t* pa, const uint8_t* sc, size_t m, int nbits, kzgb_ctx*) {
    std::vector<Fp> pts(2 * m + 2);
    std::vector<u32> k(8 * m + 8);
    for (size_t i = 0; i < m; ++i) {
        G1Aff p;
        if (!aff_from_be96(p, pa + 96 * i)) return KZGB_BADARGS;
        pts[2 * i] = p.x; pts[2 * i + 1] = p.y;
        Fr v;
        fr_raw_from_be(v, sc + 32 * i);
        if (!fr_raw_is_canonical(v)) return KZGB_BADARGS;
        for (int j = 0; j < 8; ++j) k[8 * i + j] = v.v[j];
    }
    if (m == 0) { memset(out, 0, 96); return KZGB_OK; }
    aff_to_be96(out, jac_to_aff(emu_msm(pts.data(), k.data(), 8, m, nbits)));
    return KZGB_OK;
}
kzgb_ret kzgb_pairing_check(bool* ok, const uint8_t A[96], const uint8_t B[96], kzgb_ctx* c) {
    G1Aff a, b;
    if (!aff_from_be96(a, A) || !aff_from_be96(b, B)) return KZGB_BADARGS;
    G1Jac AB[2] = {jac_from_aff(a), jac_from_aff(b)};
    PairScratch S;
    coop_pairing_check(S, c->lines, AB);
    *ok = S.result == 1;
    return KZGB_OK;
}
kzgb_ret kzgb_synth_instance(kzgb_ctx*, uint64_t seed, uint64_t offset, size_t n, uint8_t* C, uint8_t* z, uint8_t* y, uint8_t* pi, int) {
    G1Aff g = {fp_const(G1_GEN_X), fp_const(G1_GEN_Y)};
    for (size_t i = 0; i < n; ++i) {
        u64 idx = offset + i;
        Fr a = prng_fr(seed, 1, idx), zz = prng_fr(seed, 2, idx), yy = prng_fr(seed, 3, idx);
        Fr q = fr_mul(fr_sub(a, yy), fr_inv(fr_sub(fr_const(FR_TEST_TAU), zz)));
        Fr ac = fr_from_mont(a), qc = fr_from_mont(q);
        u32 w[12];
        g1_compress_words(w, jac_to_aff(jac_mul_limbs(jac_from_aff(g), ac.v, 8)));
        words_to_be(C + 48 * i, w, 12);
        g1_compress_words(w, jac_to_aff(jac_mul_limbs(jac_from_aff(g), qc.v, 8)));
        words_to_be(pi + 48 * i, w, 12);
        fr_raw_to_be(z + 32 * i, fr_from_mont(zz));
        fr_raw_to_be(y + 32 * i, fr_from_mont(yy));
    }
    return KZGB_OK;
}

static void fp12_from_be(Fp12& f, const u8* in) {
    for (int t = 0; t < 12; ++t) { Fp v; fp_from_be(v, in + 48 * t); if (t & 1) f.c[t >> 1].c1 = v; else f.c[t >> 1].c0 = v; }
}
static void fp12_to_be(u8* out, const Fp12& f) {
    for (int t = 0; t < 12; ++t) fp_to_be(out + 48 * t, (t & 1) ? f.c[t >> 1].c1 : f.c[t >> 1].c0);
}
kzgb_ret kzgb_debug_op(kzgb_ctx* c, int op, const uint8_t* in, uint8_t* out, size_t count) {
    for (size_t i = 0; i < count; ++i) {
        switch (op) {
            case 1: case 3: case 4: {
                Fp a, b;
                fp_from_be(a, in + 96 * i); fp_from_be(b, in + 96 * i + 48);
                fp_to_be(out + 48 * i, op == 1 ? fp_mul(a, b) : (op == 3 ? fp_add(a, b) : fp_sub(a, b)));
                break;
            }
            case 2: case 5: case 6: {
                Fp a;
                fp_from_be(a, in + 48 * i);
                fp_to_be(out + 48 * i, op == 2 ? fp_sqr(a) : (op == 5 ? fp_inv(a) : fp_sqrt_candidate(a)));
                break;
            }
            case 7: case 8: {
                Fr a, b;
                fr_raw_from_be(a, in + 64 * i); fr_raw_from_be(b, in + 64 * i + 32);
                a = fr_to_mont(a); b = fr_to_mont(b);
                fr_raw_to_be(out + 32 * i, fr_from_mont(op == 7 ? fr_mul(a, b) : fr_add(a, b)));
                break;
            }
            case 9: {
                G1Aff p, q;
                aff_from_be96(p, in + 192 * i); aff_from_be96(q, in + 192 * i + 96);
                aff_to_be96(out + 96 * i, jac_to_aff(jac_add(jac_from_aff(p), jac_from_aff(q))));
                break;
            }
            case 10: case 12: {
                G1Aff p;
                aff_from_be96(p, in + 96 * i);
                G1Jac r = aff_is_inf(p) ? jac_inf() : (op == 10 ? jac_dbl(jac_from_aff(p)) : jac_mul_xabs(jac_mul_xabs_aff(p)));
                aff_to_be96(out + 96 * i, jac_to_aff(r));
                break;
            }
            case 11: {
                G1Aff p;
                aff_from_be96(p, in + 128 * i);
                Fr k;
                fr_raw_from_be(k, in + 128 * i + 96);
                aff_to_be96(out + 96 * i, jac_to_aff(jac_mul_limbs(jac_from_aff(p), k.v, 8)));
                break;
            }
            case 13: case 14: case 15: case 16: case 17: case 18: {
                PairScratch S;
                if (op == 13) { fp12_from_be(S.a, in + 1152 * i); fp12_from_be(S.b, in + 1152 * i + 576); coop_mul(S, S.f, S.a, S.b); }
                else if (op == 14) { fp12_from_be(S.a, in + 576 * i); coop_frob1(S.f, S.a); }
                else if (op == 15) { fp12_from_be(S.a, in + 576 * i); coop_frob2(S.f, S.a); }
                else if (op == 16) { fp12_from_be(S.f, in + 576 * i); coop_inv(S, S.l0, S.f); coop_copy(S.f, S.l0); }
                else if (op == 17) { fp12_from_be(S.f, in + 576 * i); coop_final_exp(S); }
                else {
                    G1Aff a, b;
                    aff_from_be96(a, in + 192 * i); aff_from_be96(b, in + 192 * i + 96);
                    G1Jac P[2] = {jac_from_aff(a), jac_from_aff(b)};
                    coop_miller(S, c->lines, P);
                    coop_final_exp(S);
                }
                fp12_to_be(out + 576 * i, S.f);
                break;
            }
            case 19: {
                u32 h[8], w[16];
                sha256_init(h);
                words_from_be(w, in + 64 * i, 16);
                sha256_compress(h, w);
                for (int k = 0; k < 14; ++k) w[k] = 0;
                w[0] = 0x80000000u; w[14] = 0; w[15] = 512;
                sha256_compress(h, w);
                words_to_be(out + 32 * i, h, 8);
                break;
            }
            case 20: case 21: {
                // FP64-limb field: the h chains need round-toward-zero; every other operation is exact
                const int old = fegetround();
                fesetround(FE_TOWARDZERO);
                Fp a, b;
                if (op == 20) {
                    fp_from_be(a, in + 96 * i); fp_from_be(b, in + 96 * i + 48);
                    a = fpd_to_fp(fpd_mul(fpd_from_fp(a), fpd_from_fp(b)));
                } else {
                    fp_from_be(a, in + 48 * i);
                    FpD x = fpd_from_fp(a);
                    for (int k = 0; k < 64; ++k) x = fpd_sqr(x);
                    a = fpd_to_fp(x);
                }
                fesetround(old);
                fp_to_be(out + 48 * i, a);
                break;
            }
            default: return KZGB_BADARGS;
        }
    }
    return KZGB_OK;
}

// cell batch through the device bodies (cells.cuh), one "block" per opening
kzgb_ret verify_cell_kzg_proof_batch(bool* ok, const uint8_t* comms, size_t nc, const uint32_t* ci, const uint32_t* xi,
                                     const uint8_t* cells, const uint8_t* proofs, size_t m, kzgb_ctx* c) {
    *ok = false;
    if (!c->cell_ready || !m || !nc) return KZGB_BADARGS;
    memset(&c->art, 0, sizeof c->art);
    c->art.n = m;
    c->have_sums = false; c->have_ab = false;
    const size_t M = m + nc + 64;
    std::vector<Fp> pts(2 * M);
    u32 badp = 0, bads = 0;
    for (size_t i = 0; i < m + nc; ++i) {
        u32 w[12];
        words_from_be(w, i < m ? proofs + 48 * i : comms + 48 * (i - m), 12);
        G1Aff p;
        badp += g1_decompress_validate(p, w) != ST_OK;
        pts[2 * i] = p.x; pts[2 * i + 1] = p.y;
    }
    for (int j = 0; j < 64; ++j) { pts[2 * (m + nc + j)] = c->cell_g1[j].x; pts[2 * (m + nc + j) + 1] = c->cell_g1[j].y; }
    // leaves -> chunk digests -> root
    std::vector<u32> leaves(8 * m);
    for (size_t k = 0; k < m; ++k)
        fs_cell_leaf_words(&leaves[8 * k], ci[k], xi[k], reinterpret_cast<const u32*>(cells + 2048 * k), reinterpret_cast<const u32*>(proofs + 48 * k));
    size_t nch = (m + KZ_FS_CHUNK - 1) / KZ_FS_CHUNK;
    std::vector<u8> dig(32 * nch);
    for (size_t j = 0; j < nch; ++j) {
        u32 h[8];
        fs_chunk_words(h, &leaves[8 * j * KZ_FS_CHUNK], (u32)std::min<size_t>(KZ_FS_CHUNK, m - j * KZ_FS_CHUNK));
        words_to_be(&dig[32 * j], h, 8);
    }
    auto sha_msg = [](std::vector<u8> msg, u8 out[32]) {
        u64 bits = msg.size() * 8;
        msg.push_back(0x80);
        while (msg.size() % 64 != 56) msg.push_back(0);
        for (int i = 0; i < 8; ++i) msg.push_back((u8)(bits >> (56 - 8 * i)));
        u32 h[8];
        sha256_init(h);
        for (size_t b = 0; b < msg.size(); b += 64) { u32 w[16]; words_from_be(w, &msg[b], 16); sha256_compress(h, w); }
        words_to_be(out, h, 8);
    };
    u8 cdig[32], root[32];
    {
        std::vector<u8> msg;
        const char* tag = "KZGB200/comm_v1_";
        msg.insert(msg.end(), tag, tag + 16);
        msg.insert(msg.end(), comms, comms + 48 * nc);
        sha_msg(msg, cdig);
        std::vector<u8> r;
        const char* tag2 = "KZGB200/croot_v1";
        r.insert(r.end(), tag2, tag2 + 16);
        for (int i = 0; i < 8; ++i) r.push_back((u8)((u64)nc >> (56 - 8 * i)));
        for (int i = 0; i < 8; ++i) r.push_back((u8)((u64)m >> (56 - 8 * i)));
        r.insert(r.end(), cdig, cdig + 32);
        r.insert(r.end(), dig.begin(), dig.end());
        sha_msg(r, root);
    }
    memcpy(c->art.root, root, 32);
    u32 rw[8];
    words_from_be(rw, root, 8);
    std::vector<Fr> coefs(64 * m);
    std::vector<u32> r(4 * m), rz(8 * M);
    for (size_t k = 0; k < m; ++k) {
        CellScratch S;
        coop_cell_body(S, c->W.data(), rw, (u64)k, ci[k], xi[k], (u32)nc, cells + 2048 * k, &coefs[64 * k], &r[4 * k], &rz[8 * k]);
        bads += S.bad;
    }
    c->art.n_bad_points = badp; c->art.n_bad_scalars = bads;
    if (badp || bads) return KZGB_BADARGS;
    for (size_t i = 0; i < nc; ++i) {            // w_i
        Fr acc = fr_zero();
        for (size_t k = 0; k < m; ++k) if (ci[k] == i) { Fr v = fr_zero(); for (int j = 0; j < 4; ++j) v.v[j] = r[4 * k + j]; acc = fr_add(acc, v); }
        for (int j = 0; j < 8; ++j) rz[8 * (m + i) + j] = acc.v[j];
    }
    for (int i = 0; i < 64; ++i) {               // -S_i
        Fr acc = fr_zero();
        for (size_t k = 0; k < m; ++k) acc = fr_add(acc, coefs[64 * k + i]);
        Fr neg = fr_from_mont(fr_neg(acc));
        for (int j = 0; j < 8; ++j) rz[8 * (m + nc + i) + j] = neg.v[j];
    }
    G1Jac a = emu_msm(pts.data(), rz.data(), 8, M, 255);
    G1Jac bsum = emu_msm(pts.data(), r.data(), 4, m, 128);
    c->ab[0] = a; c->ab[1] = jac_neg(bsum);
    c->have_ab = true;
    PairScratch S;
    coop_pairing_check(S, c->lines_cell, c->ab);
    *ok = S.result == 1;
    return KZGB_OK;
}

// entry points of kzgb200.h that the emulation does not model

kzgb_ret verify_kzg_proof_batch_device(bool*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, size_t, kzgb_ctx*, void*) { return KZGB_ERROR; }
kzgb_ret kzgb_shard_phase1(kzgb_ctx*, int, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, size_t, int, void*, uint8_t*, uint32_t*) { return KZGB_ERROR; }
kzgb_ret kzgb_fs_root(uint8_t*, const uint8_t*, size_t, uint64_t) { return KZGB_ERROR; }
kzgb_ret kzgb_shard_phase2(kzgb_ctx*, int, const uint8_t*, uint64_t, void*, uint8_t*) { return KZGB_ERROR; }
kzgb_ret kzgb_combine_verify(kzgb_ctx*, const uint8_t*, int, bool*) { return KZGB_ERROR; }
kzgb_ret kzgb_shard_phase2_terms(kzgb_ctx*, int, const uint8_t*, uint64_t, void*, uint8_t*) { return KZGB_ERROR; }
kzgb_ret kzgb_shard_finish(kzgb_ctx*, int, uint32_t*, uint32_t*) { return KZGB_ERROR; }
kzgb_ret kzgb_pipeline_init(kzgb_ctx*, int) { return KZGB_ERROR; }
kzgb_ret verify_kzg_proof_batch_eip4844(bool*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, size_t, kzgb_ctx*) { return KZGB_ERROR; }
kzgb_ret verify_blob_kzg_proof_batch_eip4844(bool*, const uint8_t*, const uint8_t*, const uint8_t*, size_t, kzgb_ctx*) { return KZGB_ERROR; }
kzgb_ret kzgb_blob_challenges_evals_eip4844(uint8_t*, uint8_t*, const uint8_t*, const uint8_t*, size_t, kzgb_ctx*) { return KZGB_ERROR; }
kzgb_ret kzgb_load_trusted_setup_file(kzgb_ctx**, const char*, const int*, int, size_t) { return KZGB_ERROR; }
kzgb_ret verify_kzg_proof_batch_submit(uint64_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, size_t, int, kzgb_ctx*) { return KZGB_ERROR; }
kzgb_ret verify_kzg_proof_batch_wait(bool*, uint64_t, kzgb_ctx*) { return KZGB_ERROR; }
kzgb_ret kzgb_combine_verify_terms(kzgb_ctx*, const uint8_t*, int, bool*) { return KZGB_ERROR; }
kzgb_ret kzgb_g1_msm_times(float*, kzgb_ctx*) { return KZGB_ERROR; }
kzgb_ret kzgb_imad_peak(kzgb_ctx*, double*, double*) { return KZGB_ERROR; }
kzgb_ret kzgb_imad32_peak(kzgb_ctx*, double*, double*) { return KZGB_ERROR; }
kzgb_ret kzgb_last_stage_ms(kzgb_ctx*, float*) { return KZGB_ERROR; }
uint64_t kzgb_launch_count(const kzgb_ctx*) { return 0; }
int kzgb_set_threads(kzgb_ctx*, int) { return 0; }
// blob batch through the device bodies (blob.cuh): lanes and blobs become loops
static const std::vector<Fr>& emu_twiddles() {
    static const std::vector<Fr> W = [] {
        std::vector<Fr> w(KZ_N_EXT);
        Fr base = fr_const(FR_OMEGA_INV), acc = fr_const(FR_ONE);
        for (u32 t = 0; t < KZ_N_EXT; ++t) { w[t] = acc; acc = fr_mul(acc, base); }
        return w;
    }();
    return W;
}
static u32 emu_blob_zy(u8* z_out, u8* y_out, const u8* blobs, const u8* comms, const u8* z_in, size_t m) {
    const std::vector<Fr>& W = emu_twiddles();
    u32 bad = 0;
    for (size_t j = 0; j < m; ++j) {
        const u8* blob = blobs + (size_t)KZ_BLOB_LEN * 32 * j;
        Fr zr;
        if (z_in) fr_raw_from_be(zr, z_in + 32 * j);
        else {
            std::vector<u32> leaves(8 * KZ_BLOB_LEAVES), piece(256), cw(12);
            for (int k = 0; k < KZ_BLOB_LEAVES; ++k) { memcpy(piece.data(), blob + 1024 * k, 1024); blob_leaf_words(&leaves[8 * k], piece.data()); }
            memcpy(cw.data(), comms + 48 * j, 48);
            zr = blob_z(cw.data(), leaves.data());
        }
        const bool z_ok = fr_raw_is_canonical(zr);
        bad += !z_ok;
        if (z_out) fr_raw_to_be(z_out + 32 * j, zr);
        Fr z = fr_to_mont(z_ok ? zr : fr_zero()), total = fr_zero(), hit = fr_zero();
        bool has_hit = false;
        for (u32 lane = 0; lane < KZ_BLOB_THREADS; ++lane) {
            BlobLane L = blob_eval_lane(W.data(), blob, z, lane);
            total = fr_add(total, L.sum);
            if (L.has_hit) { has_hit = true; hit = L.hit; }
            bad += L.bad;
        }
        fr_raw_to_be(y_out + 32 * j, fr_from_mont(has_hit ? hit : blob_eval_finish(z, total)));
    }
    return bad;
}
kzgb_ret kzgb_blob_challenges_evals(uint8_t* z_out, uint8_t* y_out, const uint8_t* blobs, const uint8_t* comms, size_t m, kzgb_ctx* c) {
    if (!z_out || !y_out || !blobs || !comms || !c || m == 0) return KZGB_BADARGS;
    u32 bad = emu_blob_zy(z_out, y_out, blobs, comms, nullptr, m);
    memset(&c->art, 0, sizeof c->art);
    c->art.n = m; c->art.n_bad_scalars = bad;
    return bad ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_blob_eval(uint8_t* y_out, const uint8_t* blobs, const uint8_t* z_in, size_t m, kzgb_ctx* c) {
    if (!y_out || !blobs || !z_in || !c || m == 0) return KZGB_BADARGS;
    return emu_blob_zy(nullptr, y_out, blobs, nullptr, z_in, m) ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret verify_blob_kzg_proof_batch(bool* ok, const uint8_t* blobs, const uint8_t* comms, const uint8_t* proofs, size_t m, kzgb_ctx* c) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || !blobs || !comms || !proofs || m == 0) return KZGB_BADARGS;
    std::vector<u8> z(32 * m), y(32 * m);
    kzgb_ret rc = kzgb_blob_challenges_evals(z.data(), y.data(), blobs, comms, m, c);
    if (rc) return rc;
    return emu_verify(ok, comms, z.data(), y.data(), proofs, m, c, false);
}
kzgb_ret kzgb_set_subgroup_batch_min(kzgb_ctx* c, size_t n_min) { if (!c) return KZGB_BADARGS; c->sg_min = n_min; return KZGB_OK; }
}
