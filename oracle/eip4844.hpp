// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  EIP-4844 / c-kzg-4844 transcript mode, restated from the
// published consensus-spec functions verify_kzg_proof_batch / compute_challenge / evaluate_polynomial_in_evaluation_form
// (deneb/polynomial-commitments.md; c-kzg-4844 is absent from /root/reference, which holds only LICENSE:1-201 -- "parity
// unpinned": no c-kzg vector exists offline, the pin is tests/test_eip4844.py's pure-Python restatement).
#pragma once
#include "blobs.hpp"

namespace orc {

static const char DOM_BATCH[] = "RCKZGBATCH___V1_";
static const char DOM_BLOB[] = "FSBLOBVERIFY_V1_";

inline Fr hash_to_bls_field(const u8 d[32]) {
    u8 wide[64] = {0};
    memcpy(wide + 32, d, 32);
    return fr_from_512(wide);
}
// SHA256(domain | u64be(4096) | u64be(n) | C_i | z_i | y_i | pi_i ...)
inline void eip_batch_hash(u8 out[32], const u8* C, const u8* z, const u8* y, const u8* pi, size_t n) {
    Sha256 s;
    s.update(DOM_BATCH, 16);
    s.update_u64be(FIELD_ELEMENTS_PER_BLOB);
    s.update_u64be((u64)n);
    for (size_t i = 0; i < n; ++i) { s.update(C + 48 * i, 48); s.update(z + 32 * i, 32); s.update(y + 32 * i, 32); s.update(pi + 48 * i, 48); }
    s.final(out);
}
// compute_challenge: SHA256(domain | u128be(4096) | blob | commitment) mod r
inline Fr eip_blob_challenge(const u8* blob, const u8* C) {
    Sha256 s;
    s.update(DOM_BLOB, 16);
    s.update_u64be(0);
    s.update_u64be(FIELD_ELEMENTS_PER_BLOB);
    s.update(blob, BLOB_BYTES);
    s.update(C, 48);
    u8 d[32];
    s.final(d);
    return hash_to_bls_field(d);
}
// returns 0 ok / 1 malformed input (counts in art)
inline int verify_batch_eip4844(bool& ok, Artifacts& art, const Setup& st, const u8* C, const u8* z, const u8* y, const u8* pi, size_t n,
                                int threads) {
    ok = false;
    art = Artifacts();
    art.n = n;
    if (n == 0) return 1;
    std::vector<G1A> c(n), p(n);
    std::vector<Fr> zs(n), ys(n);
    std::atomic<u32> badp{0}, bads{0};
    parallel_for(n, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            badp += g1_decompress(c[i], C + 48 * i) != ST_OK;
            badp += g1_decompress(p[i], pi + 48 * i) != ST_OK;
            bads += !fr_from_be(zs[i], z + 32 * i);
            bads += !fr_from_be(ys[i], y + 32 * i);
        }
    });
    art.n_bad_points = badp; art.n_bad_scalars = bads;
    eip_batch_hash(art.root, C, z, y, pi, n);
    if (badp || bads) return 1;
    Fr r = hash_to_bls_field(art.root), pw = Fr::one(), sry = Fr::zero();
    std::vector<std::array<u64, 4>> rp(n), rz(n);
    for (size_t i = 0; i < n; ++i) {
        pw.to_raw(rp[i].data());
        (pw * zs[i]).to_raw(rz[i].data());
        sry = sry + pw * ys[i];
        pw = pw * r;
    }
    G1J s1 = msm(c.data(), (const u64(*)[4])rp.data(), n, 255, threads);
    G1J s2 = msm(p.data(), (const u64(*)[4])rz.data(), n, 255, threads);
    G1J s3 = msm(p.data(), (const u64(*)[4])rp.data(), n, 255, threads);
    u64 k[4];
    sry.to_raw(k);
    G1J a = s1.add(s2).add(st.g1.jac().mul(k, 4).neg());
    art.S1 = g1_affine(s1); art.S2 = g1_affine(s2); art.S3 = g1_affine(s3);
    art.A = g1_affine(a); art.B = g1_affine(s3.neg());
    art.sum_ry = sry;
    G1A P[2] = {art.A, art.B};
    G2A Q[2] = {st.g2_0, st.g2_1};
    ok = pairing_product_is_one(P, Q, 2);
    return 0;
}
// z_j, y_j of the EIP-4844 blob path; returns the number of malformed blob elements
inline unsigned blob_challenges_evals_eip4844(u8* z_out, u8* y_out, const u8* blobs, const u8* comms, size_t m, int threads) {
    std::atomic<unsigned> bad{0};
    parallel_for(m, threads, [&](size_t b, size_t e) {
        for (size_t j = b; j < e; ++j) {
            Fr zj = eip_blob_challenge(blobs + BLOB_BYTES * j, comms + 48 * j), yj;
            bad += blob_eval(yj, blobs + BLOB_BYTES * j, zj);
            zj.to_bytes_be(z_out + 32 * j);
            yj.to_bytes_be(y_out + 32 * j);
        }
    });
    return bad;
}

}  // namespace orc
