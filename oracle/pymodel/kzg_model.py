"""Pure-Python model of the KZG batch-verification semantics (docs: DESIGN.md "SPEC").  TEST INFRASTRUCTURE.

Follows BASELINE.json:5 and SURVEY.md Appendix B (the upstream reference is LICENSE-only,
/root/reference/LICENSE:1-201, so there is no reference file to cite).  Small n only.
"""
from __future__ import annotations

import hashlib
import struct

from . import bls12_381 as bls
from .bls12_381 import P, R

CHUNK = 128
FIELD_ELEMENTS_PER_BLOB = 4096

TAG_LEAF = b"KZGB200/leaf_v1_"
TAG_CHUNK = b"KZGB200/chunk_v1"
TAG_ROOT = b"KZGB200/root_v1_"
TAG_R = b"KZGB200/r_v1____"
TAG_PRNG = b"kzgb200/prng"
assert all(len(t) == 16 for t in (TAG_LEAF, TAG_CHUNK, TAG_ROOT, TAG_R))

TAU = int.from_bytes(hashlib.sha256(b"kzgb200/insecure-test-tau").digest(), "big") % R


def sha(b: bytes) -> bytes:
    return hashlib.sha256(b).digest()


# ----------------------------------------------------------------------------- PRNG
def prng_block(seed: int, stream: int, k: int) -> bytes:
    return sha(TAG_PRNG + struct.pack(">QQQ", seed, stream, k))


def prng_fr(seed: int, stream: int, idx: int) -> int:
    """idx-th Fr sample of a stream: blocks 2*idx, 2*idx+1 as a 512-bit big-endian integer mod r."""
    return int.from_bytes(prng_block(seed, stream, 2 * idx) + prng_block(seed, stream, 2 * idx + 1), "big") % R


STREAM_A, STREAM_Z, STREAM_Y, STREAM_PLANT, STREAM_POLY = 1, 2, 3, 99, 7


# ----------------------------------------------------------------------------- instances
def gen_instance_shortcut(seed: int, n: int, offset: int = 0):
    """Scalar-shortcut instances (configs BJ:8-10): a_i stands for f_i(tau)."""
    C, Z, Y, PI = [], [], [], []
    for i in range(offset, offset + n):
        a, z, y = prng_fr(seed, STREAM_A, i), prng_fr(seed, STREAM_Z, i), prng_fr(seed, STREAM_Y, i)
        q = (a - y) * pow((TAU - z) % R, R - 2, R) % R
        C.append(bls.g1_compress(bls.g1_mul(a, bls.G1)))
        PI.append(bls.g1_compress(bls.g1_mul(q, bls.G1)))
        Z.append(z.to_bytes(32, "big"))
        Y.append(y.to_bytes(32, "big"))
    return b"".join(C), b"".join(Z), b"".join(Y), b"".join(PI)


def gen_instance_poly(seed: int, n: int, degree_plus_1: int = FIELD_ELEMENTS_PER_BLOB):
    """Real-polynomial instances (config BJ:7): f_i with `degree_plus_1` random coefficients."""
    C, Z, Y, PI = [], [], [], []
    for i in range(n):
        ft, fz = 0, 0
        z = prng_fr(seed, STREAM_Z, i)
        for j in range(degree_plus_1 - 1, -1, -1):          # Horner, highest coefficient first
            c = prng_fr(seed, STREAM_POLY, i * degree_plus_1 + j)
            ft = (ft * TAU + c) % R
            fz = (fz * z + c) % R
        q = (ft - fz) * pow((TAU - z) % R, R - 2, R) % R
        C.append(bls.g1_compress(bls.g1_mul(ft, bls.G1)))
        PI.append(bls.g1_compress(bls.g1_mul(q, bls.G1)))
        Z.append(z.to_bytes(32, "big"))
        Y.append(fz.to_bytes(32, "big"))
    return b"".join(C), b"".join(Z), b"".join(Y), b"".join(PI)


def plant_index(seed: int, n: int) -> int:
    return int.from_bytes(prng_block(seed, STREAM_PLANT, 0)[:8], "big") % n


def plant_invalid(PI: bytes, j: int) -> bytes:
    """pi_j <- pi_j + G1: a valid subgroup point that is the wrong opening."""
    st, pt = bls.g1_decompress(PI[48 * j:48 * j + 48])
    assert st == 0
    return PI[:48 * j] + bls.g1_compress(bls.g1_add(pt, bls.G1)) + PI[48 * j + 48:]


def setup_g1(n1: int) -> bytes:
    return b"".join(bls.g1_compress(bls.g1_mul(pow(TAU, i, R), bls.G1)) for i in range(n1))


def setup_g2(n2: int) -> bytes:
    return b"".join(bls.g2_compress(bls.g2_mul(pow(TAU, i, R), bls.G2)) for i in range(n2))


# ----------------------------------------------------------------------------- Fiat-Shamir (App. B.4)
def fs_leaves(C, Z, Y, PI, n):
    return [sha(TAG_LEAF + C[48 * i:48 * i + 48] + Z[32 * i:32 * i + 32] + Y[32 * i:32 * i + 32] + PI[48 * i:48 * i + 48])
            for i in range(n)]


def fs_chunk_digests(leaves):
    return [sha(TAG_CHUNK + b"".join(leaves[j:j + CHUNK])) for j in range(0, len(leaves), CHUNK)]


def fs_root(digests, n):
    return sha(TAG_ROOT + struct.pack(">QQ", FIELD_ELEMENTS_PER_BLOB, n) + b"".join(digests))


def fs_r(root: bytes, i: int) -> int:
    return int.from_bytes(sha(TAG_R + root + struct.pack(">Q", i))[:16], "big")


def fs_challenges(C, Z, Y, PI, n):
    root = fs_root(fs_chunk_digests(fs_leaves(C, Z, Y, PI, n)), n)
    return root, [fs_r(root, i) for i in range(n)]


# ----------------------------------------------------------------------------- verification
KZGB_OK, KZGB_BADARGS = 0, 1


def batch_artifacts(C, Z, Y, PI, n, single=False):
    """Returns dict(ret, statuses, S1, S2, S3, A, B, sum_ry, root, r) with affine points / None."""
    out = {"ret": KZGB_OK}
    cs = [bls.g1_decompress(C[48 * i:48 * i + 48]) for i in range(n)]
    ps = [bls.g1_decompress(PI[48 * i:48 * i + 48]) for i in range(n)]
    zs = [int.from_bytes(Z[32 * i:32 * i + 32], "big") for i in range(n)]
    ys = [int.from_bytes(Y[32 * i:32 * i + 32], "big") for i in range(n)]
    out["status_c"] = [s for s, _ in cs]
    out["status_pi"] = [s for s, _ in ps]
    if any(s for s, _ in cs) or any(s for s, _ in ps) or any(v >= R for v in zs) or any(v >= R for v in ys):
        out["ret"] = KZGB_BADARGS
        return out
    if single:
        root, r = bytes(32), [1]
    else:
        root, r = fs_challenges(C, Z, Y, PI, n)
    s1 = s2 = s3 = None
    sum_ry = 0
    for i in range(n):
        s1 = bls.g1_add(s1, bls.g1_mul(r[i], cs[i][1]))
        s2 = bls.g1_add(s2, bls.g1_mul(r[i] * zs[i] % R, ps[i][1]))
        s3 = bls.g1_add(s3, bls.g1_mul(r[i], ps[i][1]))
        sum_ry = (sum_ry + r[i] * ys[i]) % R
    a = bls.g1_add(bls.g1_add(s1, s2), bls.g1_neg(bls.g1_mul(sum_ry, bls.G1)))
    b = bls.g1_neg(s3)
    out.update(root=root, r=r, S1=s1, S2=s2, S3=s3, A=a, B=b, sum_ry=sum_ry)
    return out


def verdict_tau_shortcut(art) -> bool:
    """e(A,G2) e(B,[tau]G2) = 1  <=>  A + tau*B = O  (SURVEY 4.2); pairing-free."""
    return bls.g1_add(art["A"], bls.g1_mul(TAU, art["B"])) is None


def verdict_pairing(art, g2_tau) -> bool:
    return bls.pairing_product_is_one([(art["A"], bls.G2), (art["B"], g2_tau)])


# ============================================================================= cell batch (BASELINE.json config[4])
# PeerDAS-shaped multi-point openings.  SPEC (DESIGN.md "Cell batch"): extended domain of 8192 points,
# omega = 7^((r-1)/8192); cell c (0..127) holds the 64 evaluations f(h_c * omega_64^j), j = 0..63 in natural
# order, omega_64 = omega^128, coset shift h_c = omega^brp7(c).  A proof pi opens f on the whole coset:
#   pi = [(f(tau) - I(tau)) / (tau^64 - h_c^64)] G1,   I = interpolation polynomial of the cell (deg < 64)
# Universal check for openings k with challenges r_k:
#   e(sum r_k C_{i_k} - [sum r_k I_k(tau)] G1 + sum r_k h_k^64 pi_k, G2) * e(-sum r_k pi_k, [tau^64] G2) = 1
N_EXT, N_CELLS, CELL_LEN = 8192, 128, 64
OMEGA = pow(7, (R - 1) // N_EXT, R)
OMEGA64 = pow(OMEGA, N_EXT // CELL_LEN, R)
assert pow(OMEGA, N_EXT, R) == 1 and pow(OMEGA, N_EXT // 2, R) != 1
TAG_CELL = b"KZGB200/cell_v1_"
TAG_COMM = b"KZGB200/comm_v1_"
TAG_CROOT = b"KZGB200/croot_v1"
STREAM_BLOB = 11


def brp7(c):
    return int(f"{c:07b}"[::-1], 2)


def coset_shift(c):
    return pow(OMEGA, brp7(c), R)


def cell_points(c):
    h = coset_shift(c)
    return [h * pow(OMEGA64, j, R) % R for j in range(CELL_LEN)]


def interp_coeffs(c, ys):
    """Coefficients a_0..a_63 of the degree-<64 polynomial with I(h_c w^j) = ys[j] (naive inverse DFT)."""
    h_inv = pow(coset_shift(c), R - 2, R)
    n_inv = pow(CELL_LEN, R - 2, R)
    w_inv = pow(OMEGA64, R - 2, R)
    out = []
    for i in range(CELL_LEN):
        s = sum(ys[j] * pow(w_inv, i * j, R) for j in range(CELL_LEN)) % R
        out.append(s * n_inv % R * pow(h_inv, i, R) % R)
    return out


def gen_cells(seed, n_blobs, cells_per_blob, ncoef=64 * 3):
    """n_blobs random polynomials with `ncoef` coefficients (any count <= 4096 is a valid blob polynomial);
    for each blob the first `cells_per_blob` cells (cell index = brp-free natural 0..).  Returns
    (commitments bytes, commitment_indices, cell_indices, cells bytes, proofs bytes)."""
    comms, ci, xi, cells, proofs = [], [], [], [], []
    for bi in range(n_blobs):
        coef = [prng_fr(seed, STREAM_BLOB, bi * 4096 + j) for j in range(ncoef)]
        ev = lambda x: sum(cf * pow(x, j, R) for j, cf in enumerate(coef)) % R   # noqa: E731
        ft = ev(TAU)
        comms.append(bls.g1_compress(bls.g1_mul(ft, bls.G1)))
        for c in range(cells_per_blob):
            ys = [ev(x) for x in cell_points(c)]
            a = interp_coeffs(c, ys)
            it = sum(cf * pow(TAU, j, R) for j, cf in enumerate(a)) % R
            den = (pow(TAU, CELL_LEN, R) - pow(coset_shift(c), CELL_LEN, R)) % R
            q = (ft - it) * pow(den, R - 2, R) % R
            ci.append(bi); xi.append(c)
            cells.append(b"".join(y.to_bytes(32, "big") for y in ys))
            proofs.append(bls.g1_compress(bls.g1_mul(q, bls.G1)))
    return b"".join(comms), ci, xi, b"".join(cells), b"".join(proofs)


def cell_challenges(comms, ci, xi, cells, proofs):
    m, nc = len(ci), len(comms) // 48
    leaves = [sha(TAG_CELL + struct.pack(">QQ", ci[k], xi[k]) + cells[2048 * k:2048 * k + 2048] + proofs[48 * k:48 * k + 48])
              for k in range(m)]
    cdig = sha(TAG_COMM + comms)
    root = sha(TAG_CROOT + struct.pack(">QQ", nc, m) + cdig + b"".join(fs_chunk_digests(leaves)))
    return root, [fs_r(root, k) for k in range(m)]


def cell_batch_artifacts(comms, ci, xi, cells, proofs):
    m, nc = len(ci), len(comms) // 48
    out = {"ret": KZGB_OK}
    cs = [bls.g1_decompress(comms[48 * i:48 * i + 48]) for i in range(nc)]
    ps = [bls.g1_decompress(proofs[48 * k:48 * k + 48]) for k in range(m)]
    ys = [[int.from_bytes(cells[2048 * k + 32 * j:2048 * k + 32 * j + 32], "big") for j in range(CELL_LEN)] for k in range(m)]
    if (any(s for s, _ in cs) or any(s for s, _ in ps) or any(v >= R for row in ys for v in row)
            or any(c >= N_CELLS for c in xi) or any(i >= nc for i in ci) or m == 0):
        out["ret"] = KZGB_BADARGS
        return out
    root, r = cell_challenges(comms, ci, xi, cells, proofs)
    rlc = rlp = rpi = None
    s_coef = [0] * CELL_LEN
    for k in range(m):
        rlc = bls.g1_add(rlc, bls.g1_mul(r[k], cs[ci[k]][1]))
        h64 = pow(coset_shift(xi[k]), CELL_LEN, R)
        rlp = bls.g1_add(rlp, bls.g1_mul(r[k] * h64 % R, ps[k][1]))
        rpi = bls.g1_add(rpi, bls.g1_mul(r[k], ps[k][1]))
        a = interp_coeffs(xi[k], ys[k])
        for i in range(CELL_LEN):
            s_coef[i] = (s_coef[i] + r[k] * a[i]) % R
    rli = None
    for i in range(CELL_LEN):
        rli = bls.g1_add(rli, bls.g1_mul(s_coef[i] * pow(TAU, i, R) % R, bls.G1))     # [tau^i]G1 from the setup
    a_pt = bls.g1_add(bls.g1_add(rlc, bls.g1_neg(rli)), rlp)
    out.update(root=root, r=r, A=a_pt, B=bls.g1_neg(rpi), S=s_coef)
    return out


def cell_verdict_tau_shortcut(art) -> bool:
    return bls.g1_add(art["A"], bls.g1_mul(pow(TAU, CELL_LEN, R), art["B"])) is None


# ============================================================================= EIP-4844 transcript mode
# Restated from the consensus-spec functions (deneb/polynomial-commitments.md): verify_kzg_proof_batch,
# compute_challenge, evaluate_polynomial_in_evaluation_form, hash_to_bls_field.  No c-kzg vector exists offline.
DOM_BATCH = b"RCKZGBATCH___V1_"
DOM_BLOB = b"FSBLOBVERIFY_V1_"
OMEGA_BLOB = pow(7, (R - 1) // FIELD_ELEMENTS_PER_BLOB, R)


def eip_batch_digest(C, Z, Y, PI, n) -> bytes:
    data = DOM_BATCH + FIELD_ELEMENTS_PER_BLOB.to_bytes(8, "big") + n.to_bytes(8, "big")
    for i in range(n):
        data += C[48 * i:48 * i + 48] + Z[32 * i:32 * i + 32] + Y[32 * i:32 * i + 32] + PI[48 * i:48 * i + 48]
    return sha(data)


def eip_batch_artifacts(C, Z, Y, PI, n):
    out = {"ret": KZGB_OK}
    cs = [bls.g1_decompress(C[48 * i:48 * i + 48]) for i in range(n)]
    ps = [bls.g1_decompress(PI[48 * i:48 * i + 48]) for i in range(n)]
    zs = [int.from_bytes(Z[32 * i:32 * i + 32], "big") for i in range(n)]
    ys = [int.from_bytes(Y[32 * i:32 * i + 32], "big") for i in range(n)]
    if any(s for s, _ in cs) or any(s for s, _ in ps) or any(v >= R for v in zs) or any(v >= R for v in ys):
        out["ret"] = KZGB_BADARGS
        return out
    digest = eip_batch_digest(C, Z, Y, PI, n)
    r = int.from_bytes(digest, "big") % R
    rp = [pow(r, i, R) for i in range(n)]
    proof_lincomb = proof_z_lincomb = c_minus_y_lincomb = None
    for i in range(n):
        proof_lincomb = bls.g1_add(proof_lincomb, bls.g1_mul(rp[i], ps[i][1]))
        proof_z_lincomb = bls.g1_add(proof_z_lincomb, bls.g1_mul(rp[i] * zs[i] % R, ps[i][1]))
        c_minus_y = bls.g1_add(cs[i][1], bls.g1_neg(bls.g1_mul(ys[i], bls.G1)))
        c_minus_y_lincomb = bls.g1_add(c_minus_y_lincomb, bls.g1_mul(rp[i], c_minus_y))
    a = bls.g1_add(c_minus_y_lincomb, proof_z_lincomb)          # paired with G2
    b = bls.g1_neg(proof_lincomb)                                # paired with [tau]G2
    out.update(root=digest, r=r, A=a, B=b, sum_ry=sum(x * y for x, y in zip(rp, ys)) % R)
    return out


def eip_blob_challenge(blob: bytes, commitment: bytes) -> int:
    data = DOM_BLOB + FIELD_ELEMENTS_PER_BLOB.to_bytes(16, "big") + blob + commitment
    return int.from_bytes(sha(data), "big") % R


def _brp(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2)


def eip_blob_eval(blob: bytes, z: int) -> int:
    """evaluate_polynomial_in_evaluation_form over the bit-reversed 4096-th roots of unity"""
    n = FIELD_ELEMENTS_PER_BLOB
    f = [int.from_bytes(blob[32 * i:32 * i + 32], "big") for i in range(n)]
    dom = [pow(OMEGA_BLOB, _brp(i, 12), R) for i in range(n)]
    if z in dom:
        return f[dom.index(z)]
    acc = 0
    for fi, wi in zip(f, dom):
        acc = (acc + fi * wi % R * pow((z - wi) % R, -1, R)) % R
    return acc * (pow(z, n, R) - 1) % R * pow(n, -1, R) % R
