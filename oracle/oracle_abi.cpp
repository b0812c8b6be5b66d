// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Exports the kzgb200.h C ABI on top of the scalar
// multithreaded CPU implementation, so that tests can diff the CUDA library against it stage by
// stage, and bench.py can time it as the reported CPU baseline ("cpu_baseline", kind "port").
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
// Upstream reference is LICENSE-only (/root/reference/LICENSE:1-201); spec = BASELINE.json:5.
#include <array>
#include <chrono>
#include <new>

#include <mutex>

#include "blobs.hpp"
#include "eip4844.hpp"
#include "../include/kzgb200_testing.h"

using namespace orc;

struct kzgb_ctx {
    Setup setup;
    CellSetup cells;
    int threads;
    std::vector<Shard> shards;
    Artifacts art;
    float stage_ms[KZGB_N_STAGES] = {0};
    float msm_ms[4] = {0};
    // submit / wait: the oracle has no device to overlap with -- a submitted batch is verified at once and its result
    // parked until it is collected (same ticket rules as the product: at most `depth` uncollected tickets)
    struct Parked { bool used = false; uint64_t ticket = 0; kzgb_ret rc = KZGB_OK; bool ok = false; };
    std::vector<Parked> parked;
    uint64_t next_ticket = 0;
};

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

const char* kzgb_version(void) { return "kzgb-oracle-cpu 0.1"; }

kzgb_ret kzgb_ctx_create(kzgb_ctx** out, const uint8_t* g1m, size_t n1, const uint8_t* g2m, size_t n2, const int*,
                         int n_devices, size_t) {
    if (!out || !g1m || !g2m || n1 < 1 || n2 < 2) return KZGB_BADARGS;
    kzgb_ctx* c = new (std::nothrow) kzgb_ctx();
    if (!c) return KZGB_MALLOC;
    if (g1_decompress(c->setup.g1, g1m) != ST_OK || c->setup.g1.inf || !g2_decompress(c->setup.g2_0, g2m) ||
        !g2_decompress(c->setup.g2_1, g2m + 96)) {
        delete c;
        return KZGB_BADARGS;
    }
    c->threads = (int)std::thread::hardware_concurrency();
    if (c->threads < 1) c->threads = 1;
    c->shards.resize(n_devices > 0 ? n_devices : 1);
    if (n1 >= CELL_LEN && n2 >= CELL_LEN + 1) {          // cell batch needs [tau^j]G1, j < 64 and [tau^64]G2
        c->cells.g1.resize(CELL_LEN);
        bool okc = true;
        for (size_t j = 0; j < CELL_LEN; ++j) okc = okc && g1_decompress(c->cells.g1[j], g1m + 48 * j) == ST_OK;
        okc = okc && g2_decompress(c->cells.g2_64, g2m + 96 * CELL_LEN);
        if (!okc) { delete c; return KZGB_BADARGS; }
        c->cells.g2_0 = c->setup.g2_0;
        c->cells.ready = true;
    }
    *out = c;
    return KZGB_OK;
}
void kzgb_ctx_free(kzgb_ctx* c) { delete c; }

// the oracle always checks every point on its own; the knob exists so that both libraries export the same ABI
kzgb_ret kzgb_set_subgroup_batch_min(kzgb_ctx* c, size_t) { return c ? KZGB_OK : KZGB_BADARGS; }

int kzgb_set_threads(kzgb_ctx* c, int n) {
    if (c && n > 0) c->threads = n;
    return c ? c->threads : 0;
}

// partial = A_shard | S3_shard | sum_ry_shard with A_shard = S1 + S2 - sum_ry_shard * G1 (Jacobian, canonical BE)
static void partial_to_bytes(uint8_t* out, const Partial& p, const Setup& st) {
    u64 k[4];
    p.sum_ry.to_raw(k);
    G1J a = p.s1.add(p.s2).add(st.g1.jac().mul(k, 4).neg());
    auto put = [](uint8_t* o, const G1J& j) {
        if (j.is_inf()) { memset(o, 0, 144); return; }
        j.X.to_bytes_be(o); j.Y.to_bytes_be(o + 48); j.Z.to_bytes_be(o + 96);
    };
    put(out, a);
    put(out + 144, p.s3);
    p.sum_ry.to_bytes_be(out + 288);
}

static kzgb_ret verify_impl(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                            kzgb_ctx* c, bool single) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || !C || !z || !y || !pi || n == 0) return KZGB_BADARGS;
    double t0 = now_ms();
    Shard& sh = c->shards[0];
    size_t nch = (n + CHUNK - 1) / CHUNK;
    std::vector<u8> dig(32 * nch);
    shard_phase1(sh, C, z, y, pi, n, c->threads, dig.data());
    double t1 = now_ms();
    c->art = Artifacts();
    c->art.n = n;
    c->art.n_bad_points = sh.bad_points;
    c->art.n_bad_scalars = sh.bad_scalars;
    if (sh.bad_points || sh.bad_scalars) return KZGB_BADARGS;
    if (single) memset(c->art.root, 0, 32); else fs_root(c->art.root, dig.data(), nch, n);
    Partial p = shard_phase2(sh, c->art.root, 0, c->threads, single);
    double t2 = now_ms();
    *ok = combine_verify(c->art, c->setup, &p, 1);
    double t3 = now_ms();
    memset(c->stage_ms, 0, sizeof c->stage_ms);
    c->stage_ms[1] = (float)(t1 - t0);
    c->stage_ms[6] = (float)(t2 - t1);
    c->stage_ms[8] = (float)(t3 - t2);
    c->stage_ms[9] = (float)(t3 - t0);
    return KZGB_OK;
}

kzgb_ret verify_kzg_proof(bool* ok, const uint8_t C[48], const uint8_t z[32], const uint8_t y[32], const uint8_t pi[48],
                          kzgb_ctx* c) {
    return verify_impl(ok, C, z, y, pi, 1, c, true);
}
kzgb_ret verify_kzg_proof_batch(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                                kzgb_ctx* c) {
    return verify_impl(ok, C, z, y, pi, n, c, false);
}
kzgb_ret verify_kzg_proof_batch_device(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi,
                                       size_t n, kzgb_ctx* c, void*) {
    return verify_impl(ok, C, z, y, pi, n, c, false);
}

kzgb_ret verify_kzg_proof_batch_eip4844(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                                        kzgb_ctx* c) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || !C || !z || !y || !pi || n == 0) return KZGB_BADARGS;
    bool v = false;
    int rc = verify_batch_eip4844(v, c->art, c->setup, C, z, y, pi, n, c->threads);
    *ok = v;
    return rc ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_blob_challenges_evals_eip4844(uint8_t* z_out, uint8_t* y_out, const uint8_t* blobs, const uint8_t* comms, size_t m,
                                            kzgb_ctx* c) {
    if (!z_out || !y_out || !blobs || !comms || !c || m == 0) return KZGB_BADARGS;
    return blob_challenges_evals_eip4844(z_out, y_out, blobs, comms, m, c->threads) ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret verify_blob_kzg_proof_batch_eip4844(bool* ok, const uint8_t* blobs, const uint8_t* comms, const uint8_t* proofs, size_t m,
                                             kzgb_ctx* c) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || !blobs || !comms || !proofs || m == 0) return KZGB_BADARGS;
    std::vector<u8> zs(32 * m), ys(32 * m);
    unsigned bad = blob_challenges_evals_eip4844(zs.data(), ys.data(), blobs, comms, m, c->threads);
    if (bad) { c->art = Artifacts(); c->art.n = m; c->art.n_bad_scalars = bad; return KZGB_BADARGS; }
    return verify_kzg_proof_batch_eip4844(ok, comms, zs.data(), ys.data(), proofs, m, c);
}
// c-kzg-4844 trusted_setup.txt (see kzgb200.h); the Lagrange section is skipped
kzgb_ret kzgb_load_trusted_setup_file(kzgb_ctx** out, const char* path, const int* devices, int n_devices, size_t n_max) {
    if (!out || !path) return KZGB_BADARGS;
    FILE* f = fopen(path, "r");
    if (!f) return KZGB_BADARGS;
    auto token = [&](std::vector<u8>& dst, size_t nbytes) {
        char buf[512];
        if (fscanf(f, "%400s", buf) != 1 || strlen(buf) != 2 * nbytes) return false;
        for (size_t i = 0; i < nbytes; ++i) {
            unsigned v;
            if (sscanf(buf + 2 * i, "%2x", &v) != 1) return false;
            dst.push_back((u8)v);
        }
        return true;
    };
    unsigned long n1 = 0, n2 = 0;
    bool okf = fscanf(f, "%lu %lu", &n1, &n2) == 2 && n1 >= 1 && n1 <= (1ul << 20) && n2 >= 2 && n2 <= 4096;
    std::vector<u8> skip, g1, g2;
    for (unsigned long i = 0; okf && i < n1; ++i) { skip.clear(); okf = token(skip, 48); }
    for (unsigned long i = 0; okf && i < n2; ++i) okf = token(g2, 96);
    if (okf) {
        unsigned long got = 0;
        while (got < n1 && token(g1, 48)) ++got;
        if (got == 0) {
            g1.clear();
            g1.resize(48);
            g1_compress(g1.data(), g1_generator());
        } else if (got != n1) okf = false;
    }
    fclose(f);
    if (!okf) return KZGB_BADARGS;
    return kzgb_ctx_create(out, g1.data(), g1.size() / 48, g2.data(), g2.size() / 96, devices, n_devices, n_max);
}

kzgb_ret kzgb_pipeline_init(kzgb_ctx* c, int depth) {
    if (!c || depth < 1 || depth > 8) return KZGB_BADARGS;
    for (auto& p : c->parked) if (p.used) return KZGB_BADARGS;
    c->parked.assign(depth, kzgb_ctx::Parked());
    return KZGB_OK;
}
kzgb_ret verify_kzg_proof_batch_submit(uint64_t* ticket, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi,
                                       size_t n, int, kzgb_ctx* c) {
    if (!ticket || !c || !C || !z || !y || !pi || n == 0 || c->parked.empty()) return KZGB_BADARGS;
    kzgb_ctx::Parked& p = c->parked[c->next_ticket % c->parked.size()];
    if (p.used) return KZGB_BADARGS;
    p.used = true;
    p.ticket = c->next_ticket;
    p.rc = verify_impl(&p.ok, C, z, y, pi, n, c, false);
    *ticket = c->next_ticket++;
    return KZGB_OK;
}
kzgb_ret verify_kzg_proof_batch_wait(bool* ok, uint64_t ticket, kzgb_ctx* c) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || c->parked.empty() || ticket >= c->next_ticket) return KZGB_BADARGS;
    kzgb_ctx::Parked& p = c->parked[ticket % c->parked.size()];
    if (!p.used || p.ticket != ticket) return KZGB_BADARGS;
    p.used = false;
    *ok = p.rc == KZGB_OK && p.ok;
    return p.rc;
}

kzgb_ret kzgb_shard_phase1(kzgb_ctx* c, int slot, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi,
                           size_t n, int, void*, uint8_t* dig, uint32_t* n_bad) {
    if (!c || slot < 0 || slot >= (int)c->shards.size() || !C || !z || !y || !pi || !dig || n == 0) return KZGB_BADARGS;
    Shard& sh = c->shards[slot];
    shard_phase1(sh, C, z, y, pi, n, c->threads, dig);
    if (n_bad) *n_bad = sh.bad_points + sh.bad_scalars;
    return (sh.bad_points || sh.bad_scalars) ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_fs_root(uint8_t root[32], const uint8_t* dig, size_t nch, uint64_t n_total) {
    if (!root || !dig) return KZGB_BADARGS;
    fs_root(root, dig, nch, n_total);
    return KZGB_OK;
}
kzgb_ret kzgb_shard_phase2(kzgb_ctx* c, int slot, const uint8_t root[32], uint64_t off, void*, uint8_t out[KZGB_PARTIAL_BYTES]) {
    if (!c || slot < 0 || slot >= (int)c->shards.size() || !root || !out) return KZGB_BADARGS;
    Partial p = shard_phase2(c->shards[slot], root, off, c->threads);
    partial_to_bytes(out, p, c->setup);
    return KZGB_OK;
}
kzgb_ret kzgb_combine_verify(kzgb_ctx* c, const uint8_t* parts, int np, bool* ok) {
    if (!c || !parts || np < 1 || !ok) return KZGB_BADARGS;
    *ok = false;
    G1J a = G1J::inf(), s3 = G1J::inf();
    Fr sry = Fr::zero();
    for (int i = 0; i < np; ++i) {
        const uint8_t* b = parts + (size_t)KZGB_PARTIAL_BYTES * i;
        auto get = [](G1J& j, const uint8_t* o) {
            return Fp::from_bytes_be(j.X, o) && Fp::from_bytes_be(j.Y, o + 48) && Fp::from_bytes_be(j.Z, o + 96);
        };
        G1J pa, p3;
        Fr v;
        if (!get(pa, b) || !get(p3, b + 144) || !Fr::from_bytes_be(v, b + 288)) return KZGB_BADARGS;
        a = a.add(pa); s3 = s3.add(p3); sry = sry + v;
    }
    c->art = Artifacts();
    c->art.S1 = c->art.S2 = c->art.S3 = G1A::infinity();     // per-sum artefacts are not carried by partials
    c->art.A = g1_affine(a);
    c->art.B = g1_affine(s3.neg());
    c->art.sum_ry = sry;
    G1A P[2] = {c->art.A, c->art.B};
    G2A Q[2] = {c->setup.g2_0, c->setup.g2_1};
    *ok = pairing_product_is_one(P, Q, 2);
    return KZGB_OK;
}

// Terms exchange (kzgb200.h KZGB_TERMS_BYTES): any 2 x 33 points with sum_t 2^(4t) V_t = A_shard resp. B_shard are a valid
// record.  The oracle has no bucket slices to read terms from, so it sends V_0 = A_shard, V_33 = -S3_shard and
// infinity elsewhere; its combine accepts arbitrary records (the CUDA library's included) and evaluates
// A = sum_t 2^(4t) sum_g V_t^g by Horner before the ordinary two-pairing check.
kzgb_ret kzgb_shard_phase2_terms(kzgb_ctx* c, int slot, const uint8_t root[32], uint64_t off, void*, uint8_t out[KZGB_TERMS_BYTES]) {
    if (!c || slot < 0 || slot >= (int)c->shards.size() || !root || !out) return KZGB_BADARGS;
    Partial p = shard_phase2(c->shards[slot], root, off, c->threads);
    u64 k[4];
    p.sum_ry.to_raw(k);
    G1J a = p.s1.add(p.s2).add(c->setup.g1.jac().mul(k, 4).neg());
    memset(out, 0, KZGB_TERMS_BYTES);
    auto put = [](uint8_t* o, const G1J& j) {                 // XYZZ of a Jacobian point: (X, Y, Z^2, Z^3)
        if (j.is_inf()) return;
        Fp zz = j.Z.sqr();
        j.X.to_bytes_be(o); j.Y.to_bytes_be(o + 48); zz.to_bytes_be(o + 96); (zz * j.Z).to_bytes_be(o + 144);
    };
    put(out, a);
    put(out + 192 * (KZGB_N_TERMS / 2), p.s3.neg());
    p.sum_ry.to_bytes_be(out + 192 * KZGB_N_TERMS);
    return KZGB_OK;
}
kzgb_ret kzgb_shard_finish(kzgb_ctx* c, int slot, uint32_t* bad_points, uint32_t* bad_scalars) {
    if (!c || slot < 0 || slot >= (int)c->shards.size()) return KZGB_BADARGS;
    const Shard& sh = c->shards[slot];
    if (bad_points) *bad_points = sh.bad_points;
    if (bad_scalars) *bad_scalars = sh.bad_scalars;
    return (sh.bad_points || sh.bad_scalars) ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_combine_verify_terms(kzgb_ctx* c, const uint8_t* terms, int ns, bool* ok) {
    if (!c || !terms || ns < 1 || ns > 64 || !ok) return KZGB_BADARGS;
    *ok = false;
    const int T = KZGB_N_TERMS / 2;
    G1J side[2] = {G1J::inf(), G1J::inf()};
    Fr sry = Fr::zero();
    for (int sd = 0; sd < 2; ++sd)
        for (int t = T - 1; t >= 0; --t) {
            for (int u = 0; u < 4; ++u) side[sd] = side[sd].dbl();
            for (int g = 0; g < ns; ++g) {
                const uint8_t* b = terms + (size_t)KZGB_TERMS_BYTES * g + 192 * (sd * T + t);
                Fp X, Y, ZZ, ZZZ;
                if (!Fp::from_bytes_be(X, b) || !Fp::from_bytes_be(Y, b + 48) || !Fp::from_bytes_be(ZZ, b + 96) ||
                    !Fp::from_bytes_be(ZZZ, b + 144))
                    return KZGB_BADARGS;
                if (ZZ.is_zero()) continue;
                side[sd] = side[sd].add(G1J{X * ZZ, Y * ZZZ, ZZ});       // (X/ZZ, Y/ZZZ) with Z = ZZ, as ZZ^3 = ZZZ^2
            }
        }
    for (int g = 0; g < ns; ++g) {
        Fr v;
        if (!Fr::from_bytes_be(v, terms + (size_t)KZGB_TERMS_BYTES * g + 192 * KZGB_N_TERMS)) return KZGB_BADARGS;
        sry = sry + v;
    }
    c->art = Artifacts();
    c->art.S1 = c->art.S2 = c->art.S3 = G1A::infinity();
    c->art.A = g1_affine(side[0]);
    c->art.B = g1_affine(side[1]);
    c->art.sum_ry = sry;
    G1A P[2] = {c->art.A, c->art.B};
    G2A Q[2] = {c->setup.g2_0, c->setup.g2_1};
    *ok = pairing_product_is_one(P, Q, 2);
    return KZGB_OK;
}

kzgb_ret verify_cell_kzg_proof_batch(bool* ok, const uint8_t* comms, size_t nc, const uint32_t* ci, const uint32_t* xi,
                                     const uint8_t* cells, const uint8_t* proofs, size_t m, kzgb_ctx* c) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || !comms || !ci || !xi || !cells || !proofs) return KZGB_BADARGS;
    bool v = false;
    int rc = verify_cells(v, c->art, c->cells, comms, nc, ci, xi, cells, proofs, m, c->threads);
    *ok = v;
    return rc ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_blob_challenges_evals(uint8_t* z_out, uint8_t* y_out, const uint8_t* blobs, const uint8_t* comms, size_t m, kzgb_ctx* c) {
    if (!z_out || !y_out || !blobs || !comms || !c || m == 0) return KZGB_BADARGS;
    unsigned bad = blob_challenges_evals(z_out, y_out, blobs, comms, m, c->threads);
    c->art = Artifacts();
    c->art.n = m;
    c->art.n_bad_scalars = bad;
    return bad ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_blob_eval(uint8_t* y_out, const uint8_t* blobs, const uint8_t* z_in, size_t m, kzgb_ctx* c) {
    if (!y_out || !blobs || !z_in || !c || m == 0) return KZGB_BADARGS;
    unsigned bad = 0;
    for (size_t j = 0; j < m; ++j) {
        Fr z, y;
        if (!fr_from_be(z, z_in + 32 * j)) { ++bad; continue; }
        bad += blob_eval(y, blobs + BLOB_BYTES * j, z);
        y.to_bytes_be(y_out + 32 * j);
    }
    return bad ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret verify_blob_kzg_proof_batch(bool* ok, const uint8_t* blobs, const uint8_t* comms, const uint8_t* proofs, size_t m, kzgb_ctx* c) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!c || !blobs || !comms || !proofs || m == 0) return KZGB_BADARGS;
    std::vector<u8> z(32 * m), y(32 * m);
    kzgb_ret rc = kzgb_blob_challenges_evals(z.data(), y.data(), blobs, comms, m, c);
    if (rc) return rc;
    return verify_impl(ok, comms, z.data(), y.data(), proofs, m, c, false);
}
// oracle-only: synthetic blobs with their commitments and proofs (known test tau)
kzgb_ret kzgb_oracle_synth_blobs(uint64_t seed, size_t m, uint8_t* blobs, uint8_t* comms, uint8_t* proofs, int threads) {
    if (!blobs || !comms || !proofs) return KZGB_BADARGS;
    synth_blobs(seed, m, blobs, comms, proofs, threads > 0 ? threads : (int)std::thread::hardware_concurrency());
    return KZGB_OK;
}
// oracle-only: synthetic cells (n_blobs * cells_per_blob openings)
kzgb_ret kzgb_oracle_synth_cells(uint64_t seed, size_t n_blobs, size_t cells_per_blob, size_t ncoef, uint8_t* comms, uint32_t* ci,
                                 uint32_t* xi, uint8_t* cells, uint8_t* proofs, int threads) {
    if (ncoef > 4096 || cells_per_blob > N_CELLS) return KZGB_BADARGS;
    synth_cells(seed, n_blobs, cells_per_blob, ncoef, comms, ci, xi, cells, proofs,
                threads > 0 ? threads : (int)std::thread::hardware_concurrency());
    return KZGB_OK;
}

kzgb_ret kzgb_g1_decompress_batch(uint8_t* aff, uint8_t* st, const uint8_t* in, size_t m, kzgb_ctx* c) {
    if (!aff || !st || !in || !c) return KZGB_BADARGS;
    parallel_for(m, c->threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            G1A p;
            st[i] = g1_decompress(p, in + 48 * i);
            g1_affine_bytes(aff + 96 * i, p);
        }
    });
    return KZGB_OK;
}
kzgb_ret kzgb_fs_challenges(uint8_t root[32], uint8_t* r_out, const uint8_t* C, const uint8_t* z, const uint8_t* y,
                            const uint8_t* pi, size_t n, kzgb_ctx* c) {
    if (!root || !r_out || !C || !z || !y || !pi || !c || n == 0) return KZGB_BADARGS;
    size_t nch = (n + CHUNK - 1) / CHUNK;
    std::vector<u8> dig(32 * nch);
    fs_chunk_digests(dig.data(), C, z, y, pi, n, c->threads);
    fs_root(root, dig.data(), nch, n);
    parallel_for(n, c->threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) fs_r(r_out + 16 * i, root, i);
    });
    return KZGB_OK;
}
kzgb_ret kzgb_g1_msm(uint8_t out[96], const uint8_t* pts, const uint8_t* sc, size_t m, int nbits, kzgb_ctx* c) {
    if (!out || !pts || !sc || !c || (nbits != 255 && nbits != 128)) return KZGB_BADARGS;
    std::vector<G1A> p(m);
    std::vector<std::array<u64, 4>> k(m);
    for (size_t i = 0; i < m; ++i) {
        Fr s;
        if (!g1_from_affine_bytes(p[i], pts + 96 * i) || !Fr::from_bytes_be(s, sc + 32 * i)) return KZGB_BADARGS;
        s.to_raw(k[i].data());
        if (nbits == 128 && (k[i][2] | k[i][3])) return KZGB_BADARGS;
    }
    double t0 = now_ms();
    G1J r = msm(p.data(), (const u64(*)[4])k.data(), m, nbits, c->threads);
    c->msm_ms[3] = (float)(now_ms() - t0);
    g1_affine_bytes(out, g1_affine(r));
    return KZGB_OK;
}
kzgb_ret kzgb_g1_msm_times(float ms[4], kzgb_ctx* c) {
    if (!ms || !c) return KZGB_BADARGS;
    memcpy(ms, c->msm_ms, sizeof c->msm_ms);
    return KZGB_OK;
}
kzgb_ret kzgb_pairing_check(bool* ok, const uint8_t A[96], const uint8_t B[96], kzgb_ctx* c) {
    if (!ok || !A || !B || !c) return KZGB_BADARGS;
    G1A P[2];
    if (!g1_from_affine_bytes(P[0], A) || !g1_from_affine_bytes(P[1], B)) return KZGB_BADARGS;
    G2A Q[2] = {c->setup.g2_0, c->setup.g2_1};
    *ok = pairing_product_is_one(P, Q, 2);
    return KZGB_OK;
}
kzgb_ret kzgb_last_artifacts(kzgb_ctx* c, kzgb_artifacts* o) {
    if (!c || !o) return KZGB_BADARGS;
    memset(o, 0, sizeof *o);
    g1_affine_bytes(o->S1, c->art.S1); g1_affine_bytes(o->S2, c->art.S2); g1_affine_bytes(o->S3, c->art.S3);
    g1_affine_bytes(o->A, c->art.A); g1_affine_bytes(o->B, c->art.B);
    c->art.sum_ry.to_bytes_be(o->sum_ry);
    memcpy(o->root, c->art.root, 32);
    o->n = c->art.n;
    o->n_bad_points = c->art.n_bad_points;
    o->n_bad_scalars = c->art.n_bad_scalars;
    memcpy(o->stage_ms, c->stage_ms, sizeof c->stage_ms);
    return KZGB_OK;
}

kzgb_ret kzgb_synth_instance(kzgb_ctx* c, uint64_t seed, uint64_t off, size_t n, uint8_t* C, uint8_t* z, uint8_t* y,
                             uint8_t* pi, int) {
    if (!C || !z || !y || !pi) return KZGB_BADARGS;
    synth_instance(seed, off, n, C, z, y, pi, c ? c->threads : (int)std::thread::hardware_concurrency());
    return KZGB_OK;
}
kzgb_ret kzgb_synth_setup(uint8_t* g1, size_t n1, uint8_t* g2, size_t n2) {
    if ((n1 && !g1) || (n2 && !g2)) return KZGB_BADARGS;
    synth_setup(g1, n1, g2, n2);
    return KZGB_OK;
}
// oracle-only: real-polynomial instances (config BJ:7) and planted-invalid helper
kzgb_ret kzgb_oracle_synth_instance_poly(uint64_t seed, size_t n, size_t ncoef, uint8_t* C, uint8_t* z, uint8_t* y,
                                         uint8_t* pi, int threads) {
    synth_instance_poly(seed, n, ncoef, C, z, y, pi, threads > 0 ? threads : (int)std::thread::hardware_concurrency());
    return KZGB_OK;
}
// pi_j <- pi_j + G1 in place (valid subgroup point, wrong opening)
kzgb_ret kzgb_oracle_plant_invalid(uint8_t* pi, size_t j) {
    G1A p;
    if (g1_decompress(p, pi + 48 * j) != ST_OK) return KZGB_BADARGS;
    g1_compress(pi + 48 * j, g1_affine(p.jac().add(g1_generator().jac())));
    return KZGB_OK;
}
uint64_t kzgb_oracle_plant_index(uint64_t seed, uint64_t n) {
    u8 b[32];
    prng_block(b, seed, STREAM_PLANT, 0);
    u64 v = 0;
    for (int i = 0; i < 8; ++i) v = v << 8 | b[i];
    return v % n;
}
// oracle-only: slow subgroup decision ([r]P == O) for cross-checking the fast test
int kzgb_oracle_g1_status_slow(const uint8_t in[48]) {
    G1A p;
    return g1_decompress(p, in, 2);
}
// oracle-only: tau-shortcut verdict A + tau*B == O (pairing-free)
int kzgb_oracle_tau_shortcut(const uint8_t A[96], const uint8_t B[96]) {
    G1A a, b;
    if (!g1_from_affine_bytes(a, A) || !g1_from_affine_bytes(b, B)) return -1;
    u64 k[4];
    test_tau().to_raw(k);
    return a.jac().add(b.jac().mul(k, 4)).is_inf() ? 1 : 0;
}

static void fp12_from_bytes(Fp12& f, const u8* b) {
    for (int k = 0; k < 6; ++k) {
        Fp::from_bytes_be(f.coef(k).c0, b + 96 * k);
        Fp::from_bytes_be(f.coef(k).c1, b + 96 * k + 48);
    }
}
static void fp12_to_bytes(u8* b, const Fp12& f) {
    for (int k = 0; k < 6; ++k) {
        f.coef(k).c0.to_bytes_be(b + 96 * k);
        f.coef(k).c1.to_bytes_be(b + 96 * k + 48);
    }
}

kzgb_ret kzgb_debug_op(kzgb_ctx* c, int op, const uint8_t* in, uint8_t* out, size_t count) {
    if (!in || !out) return KZGB_BADARGS;
    for (size_t i = 0; i < count; ++i) {
        switch (op) {
            case KZGB_OP_FP_MUL: case KZGB_OP_FP_ADD: case KZGB_OP_FP_SUB: {
                Fp a, b;
                if (!Fp::from_bytes_be(a, in + 96 * i) || !Fp::from_bytes_be(b, in + 96 * i + 48)) return KZGB_BADARGS;
                Fp r = op == KZGB_OP_FP_MUL ? a * b : (op == KZGB_OP_FP_ADD ? a + b : a - b);
                r.to_bytes_be(out + 48 * i);
                break;
            }
            case KZGB_OP_FP_SQR: case KZGB_OP_FP_INV: case KZGB_OP_FP_SQRT_CAND: {
                Fp a;
                if (!Fp::from_bytes_be(a, in + 48 * i)) return KZGB_BADARGS;
                Fp r;
                if (op == KZGB_OP_FP_SQR) r = a.sqr();
                else if (op == KZGB_OP_FP_INV) r = a.inv();
                else {
                    u64 e[6], onev[6] = {1};
                    add_raw<6>(e, fp_params().mod, onev);
                    for (int k = 0; k < 6; ++k) e[k] = (e[k] >> 2) | (k < 5 ? e[k + 1] << 62 : 0);
                    r = a.pow(e, 6);
                }
                r.to_bytes_be(out + 48 * i);
                break;
            }
            case KZGB_OP_FR_MUL: case KZGB_OP_FR_ADD: {
                Fr a, b;
                if (!Fr::from_bytes_be(a, in + 64 * i) || !Fr::from_bytes_be(b, in + 64 * i + 32)) return KZGB_BADARGS;
                (op == KZGB_OP_FR_MUL ? a * b : a + b).to_bytes_be(out + 32 * i);
                break;
            }
            case KZGB_OP_G1_ADD: {
                G1A p, q;
                if (!g1_from_affine_bytes(p, in + 192 * i) || !g1_from_affine_bytes(q, in + 192 * i + 96)) return KZGB_BADARGS;
                g1_affine_bytes(out + 96 * i, g1_affine(p.jac().add(q.jac())));
                break;
            }
            case KZGB_OP_G1_DBL: case KZGB_OP_G1_MUL_XSQ: {
                G1A p;
                if (!g1_from_affine_bytes(p, in + 96 * i)) return KZGB_BADARGS;
                u64 k[1] = {X_ABS};
                G1J r = op == KZGB_OP_G1_DBL ? p.jac().dbl() : p.jac().mul(k, 1).mul(k, 1);
                g1_affine_bytes(out + 96 * i, g1_affine(r));
                break;
            }
            case KZGB_OP_G1_MUL: {
                G1A p;
                Fr s;
                if (!g1_from_affine_bytes(p, in + 128 * i) || !Fr::from_bytes_be(s, in + 128 * i + 96)) return KZGB_BADARGS;
                u64 k[4];
                s.to_raw(k);
                g1_affine_bytes(out + 96 * i, g1_affine(p.jac().mul(k, 4)));
                break;
            }
            case KZGB_OP_FP12_MUL: {
                Fp12 a, b;
                fp12_from_bytes(a, in + 1152 * i);
                fp12_from_bytes(b, in + 1152 * i + 576);
                fp12_to_bytes(out + 576 * i, a * b);
                break;
            }
            case KZGB_OP_FP12_FROB1: case KZGB_OP_FP12_FROB2: case KZGB_OP_FP12_INV: case KZGB_OP_FINAL_EXP: {
                Fp12 a;
                fp12_from_bytes(a, in + 576 * i);
                Fp12 r = op == KZGB_OP_FP12_FROB1 ? frob1(a) : op == KZGB_OP_FP12_FROB2 ? frob2(a)
                         : op == KZGB_OP_FP12_INV ? a.inv() : final_exp(a);
                fp12_to_bytes(out + 576 * i, r);
                break;
            }
            case KZGB_OP_MILLER_FE: {
                if (!c) return KZGB_BADARGS;
                G1A P[2];
                if (!g1_from_affine_bytes(P[0], in + 192 * i) || !g1_from_affine_bytes(P[1], in + 192 * i + 96)) return KZGB_BADARGS;
                G2A Q[2] = {c->setup.g2_0, c->setup.g2_1};
                fp12_to_bytes(out + 576 * i, final_exp(miller_loop_multi(P, Q, 2)));
                break;
            }
            case KZGB_OP_SHA256_64: {
                Sha256 s;
                s.update(in + 64 * i, 64);
                s.final(out + 32 * i);
                break;
            }
            default: return KZGB_BADARGS;
        }
    }
    return KZGB_OK;
}

kzgb_ret kzgb_imad_peak(kzgb_ctx*, double*, double*) { return KZGB_ERROR; }
kzgb_ret kzgb_imad32_peak(kzgb_ctx*, double*, double*) { return KZGB_ERROR; }
kzgb_ret kzgb_last_stage_ms(kzgb_ctx* c, float ms[KZGB_N_STAGES]) {
    if (!c || !ms) return KZGB_BADARGS;
    memcpy(ms, c->stage_ms, sizeof c->stage_ms);
    return KZGB_OK;
}
uint64_t kzgb_launch_count(const kzgb_ctx*) { return 0; }

}  // extern "C"
