"""Parity checks shared by the emulation tests (CPU, -m "not gpu") and the CUDA tests (-m gpu).

Every check drives the library under test and the CPU oracle through the same C ABI
(include/kzgb200.h) on the same seeded inputs and demands byte equality (BASELINE.json:5: "bit-exact
against the CPU oracle: the accept/reject decision, every canonical affine MSM output and every
decompressed point").
"""
import ctypes as C
import hashlib
import random

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k
from tests.helpers import (P, R, edge_fps, f12_bytes, fp_bytes, fr_bytes, negative_g1_encodings, rand_curve_point, rand_g1)


def _same(ctx, oracle, op, data):
    rc1, o1 = ctx.debug_op(op, data)
    rc2, o2 = oracle.debug_op(op, data)
    assert rc1 == 0 and rc2 == 0, (op, rc1, rc2)
    assert o1 == o2, f"debug op {op} differs"


def check_field_ops(ctx, oracle, n_random=200):
    rnd = random.Random(101)
    vals = edge_fps() + [rnd.randrange(P) for _ in range(30)]
    pairs = [(x, y) for x in vals[:13] for y in vals[:13]] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(n_random)]
    data = b"".join(fp_bytes(x) + fp_bytes(y) for x, y in pairs)
    for op in ("FP_MUL", "FP_ADD", "FP_SUB"):
        _same(ctx, oracle, op, data)
    data = b"".join(fp_bytes(x) for x in vals)
    for op in ("FP_SQR", "FP_INV", "FP_SQRT_CAND"):
        _same(ctx, oracle, op, data)
    frs = [0, 1, R - 1, R - 2, 2 ** 128 - 1, 2 ** 254, 2 ** 32 - 1] + [rnd.randrange(R) for _ in range(60)]
    prs = [(x, y) for x in frs[:7] for y in frs[:7]] + [(rnd.randrange(R), rnd.randrange(R)) for _ in range(n_random)]
    data = b"".join(fr_bytes(x) + fr_bytes(y) for x, y in prs)
    for op in ("FR_MUL", "FR_ADD"):
        _same(ctx, oracle, op, data)


def check_fpd_ops(ctx, oracle):
    """FP64-limb multiplier (tools/microbench/fpd.cuh; emulation only): same bits as the integer Montgomery product, lazy [0, 2p) chains included."""
    rnd = random.Random(131)
    vals = edge_fps() + [(1 << 48) - 1, 1 << 48, (1 << 96) - 1, (1 << 336) - 1, P - (1 << 48)] + [rnd.randrange(P) for _ in range(60)]
    pairs = [(x, y) for x in vals[:18] for y in vals[:18]] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(400)]
    data = b"".join(fp_bytes(x) + fp_bytes(y) for x, y in pairs)
    rc, got = ctx.debug_op("FPD_MUL", data)
    rc2, want = oracle.debug_op("FP_MUL", data)
    assert rc == 0 and rc2 == 0 and got == want
    rc, got = ctx.debug_op("FPD_SQR_CHAIN", b"".join(fp_bytes(x) for x in vals))
    assert rc == 0 and got == b"".join(fp_bytes(pow(x, 1 << 64, P)) for x in vals)


def check_g1_ops(ctx, oracle):
    rnd = random.Random(102)
    pts = [rand_g1(rnd) for _ in range(5)] + [None, rand_curve_point(rnd), (0, 2)]
    pairs = [(p, q) for p in pts for q in pts] + [(pts[0], b.g1_neg(pts[0]))]
    _same(ctx, oracle, "G1_ADD", b"".join(b.g1_affine_bytes(p) + b.g1_affine_bytes(q) for p, q in pairs))
    data = b"".join(b.g1_affine_bytes(p) for p in pts)
    _same(ctx, oracle, "G1_DBL", data)
    _same(ctx, oracle, "G1_MUL_XSQ", data)
    ks = [0, 1, 2, R - 1, 2 ** 128 - 1] + [rnd.randrange(R) for _ in range(3)]
    _same(ctx, oracle, "G1_MUL", b"".join(b.g1_affine_bytes(pts[i % 5]) + fr_bytes(kk) for i, kk in enumerate(ks)))


def check_decompress(ctx, oracle, n_valid=40):
    rnd = random.Random(103)
    cases = negative_g1_encodings(rnd)
    valid = [b.g1_compress(rand_g1(rnd)) for _ in range(n_valid)]
    order = [c for c, _ in cases] + valid
    rnd.shuffle(order)
    data = b"".join(order)
    rc1, a1, s1 = ctx.g1_decompress_batch(data)
    rc2, a2, s2 = oracle.g1_decompress_batch(data)
    assert rc1 == 0 and rc2 == 0
    assert s1 == s2, "status bytes differ"
    assert a1 == a2, "decompressed points differ"
    want = dict(cases)
    for i, enc in enumerate(order):
        if enc in want:
            assert s1[i] == want[enc]
    for m in (1, 2, 3):                                   # odd / tiny sizes
        assert ctx.g1_decompress_batch(data[:48 * m]) == oracle.g1_decompress_batch(data[:48 * m])


def check_tower_and_pairing(ctx, oracle):
    rnd = random.Random(104)
    a = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(6)]
    c = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(6)]
    _same(ctx, oracle, "FP12_MUL", f12_bytes(a) + f12_bytes(c))
    sparse = [a[0], (0, 0), a[2], (a[3][0], 0), (0, 0), (0, 0)]
    _same(ctx, oracle, "FP12_MUL", f12_bytes(c) + f12_bytes(sparse))
    for op in ("FP12_FROB1", "FP12_FROB2", "FP12_INV", "FINAL_EXP"):
        _same(ctx, oracle, op, f12_bytes(a))
    A, B = rand_g1(rnd), rand_g1(rnd)
    _same(ctx, oracle, "MILLER_FE", b.g1_affine_bytes(A) + b.g1_affine_bytes(B))
    _same(ctx, oracle, "MILLER_FE", b.g1_affine_bytes(A) + bytes(96))
    Bp = rand_g1(rnd)
    Ap = b.g1_neg(b.g1_mul(k.TAU, Bp))
    for lib in (ctx, oracle):
        assert lib.pairing_check(b.g1_affine_bytes(Ap), b.g1_affine_bytes(Bp)) == (0, True)
        assert lib.pairing_check(b.g1_affine_bytes(A), b.g1_affine_bytes(B)) == (0, False)
        assert lib.pairing_check(bytes(96), bytes(96)) == (0, True)
        assert lib.pairing_check(b.g1_affine_bytes(A), bytes(96)) == (0, False)


def check_fs(ctx, oracle, sizes=(1, 2, 63, 127, 128, 129, 1023, 1024, 1025, 2500)):
    seed = 0x4B5A4703
    nmax = max(sizes)
    C, Z, Y, PI = oracle.synth_instance(seed, 0, nmax)
    for n in sizes:
        args = (C[:48 * n], Z[:32 * n], Y[:32 * n], PI[:48 * n], n)
        r1 = ctx.fs_challenges(*args)
        r2 = oracle.fs_challenges(*args)
        assert r1 == r2, f"fs challenges differ at n={n}"


def check_msm(ctx, oracle, sizes=((1, 255), (2, 255), (7, 255), (64, 255), (64, 128), (300, 255), (300, 128))):
    rnd = random.Random(105)
    base = [rand_g1(rnd) for _ in range(24)]
    for m, nbits in sizes:
        pts = [rnd.choice(base) if rnd.random() < 0.9 else None for _ in range(m)]
        ks = [rnd.randrange(2 ** 128 if nbits == 128 else R) for _ in range(m)]
        if m >= 7:
            pts[1] = pts[0]; pts[2] = b.g1_neg(pts[0])                     # duplicate and negation share buckets
            ks[1] = ks[0]; ks[2] = ks[0]; ks[3] = 0
            ks[4] = (2 ** 128 - 1) if nbits == 128 else R - 1              # all-ones / maximal scalar
            ks[5] = 1
        pb = b"".join(b.g1_affine_bytes(p_) for p_ in pts)
        kb = b"".join(fr_bytes(k_) for k_ in ks)
        r1 = ctx.g1_msm(pb, kb, nbits)
        r2 = oracle.g1_msm(pb, kb, nbits)
        assert r1[0] == 0 and r1 == r2, f"msm differs at m={m} nbits={nbits}"


def check_synth(ctx, oracle, n=10):
    seed = 0x4B5A4701
    assert ctx.synth_instance(seed, 0, n) == oracle.synth_instance(seed, 0, n)
    assert ctx.synth_instance(seed, 5, 3) == oracle.synth_instance(seed, 5, 3)


def check_verify(ctx, oracle, oracle_lib, sizes=(1, 2, 9, 64), seed=0x4B5A4701):
    """Full batch verification: verdicts, artefacts, planted invalid proof, malformed inputs."""
    import ctypes
    for n in sizes:
        C, Z, Y, PI = oracle.synth_instance(seed, 0, n)
        r1 = ctx.verify_kzg_proof_batch(C, Z, Y, PI, n)
        r2 = oracle.verify_kzg_proof_batch(C, Z, Y, PI, n)
        assert r1 == r2 == (0, True), (n, r1, r2)
        a1, a2 = ctx.last_artifacts(), oracle.last_artifacts()
        for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
            assert a1[key] == a2[key], (n, key)
        j = oracle_lib.lib.kzgb_oracle_plant_index(ctypes.c_uint64(seed), ctypes.c_uint64(n))
        bad = bytearray(PI)
        oracle_lib.lib.kzgb_oracle_plant_invalid((ctypes.c_uint8 * len(bad)).from_buffer(bad), ctypes.c_size_t(j))
        bad = bytes(bad)
        assert ctx.verify_kzg_proof_batch(C, Z, Y, bad, n) == (0, False)
        assert oracle.verify_kzg_proof_batch(C, Z, Y, bad, n) == (0, False)
        a1, a2 = ctx.last_artifacts(), oracle.last_artifacts()
        assert a1["A"] == a2["A"] and a1["B"] == a2["B"]
    C, Z, Y, PI = oracle.synth_instance(seed, 0, 4)
    for lib in (ctx, oracle):
        assert lib.verify_kzg_proof(C[:48], Z[:32], Y[:32], PI[:48]) == (0, True)
        assert lib.verify_kzg_proof(C[:48], Z[:32], Y[:32], PI[48:96]) == (0, False)
        assert lib.verify_kzg_proof_batch(b.g1_compress((0, 2)) + C[48:], Z, Y, PI, 4) == (1, False)       # off-subgroup
        assert lib.verify_kzg_proof_batch(C, R.to_bytes(32, "big") + Z[32:], Y, PI, 4) == (1, False)       # z >= r
        assert lib.verify_kzg_proof_batch(C, Z, Y[:96] + (2 ** 256 - 1).to_bytes(32, "big"), PI, 4) == (1, False)
        assert lib.verify_kzg_proof_batch(C, Z, Y, PI, 0) == (1, False)
        # infinity commitment/proof are valid encodings: C = O, pi = O verifies iff y = 0
        inf = b.g1_compress(None)
        assert lib.verify_kzg_proof(inf, Z[:32], bytes(32), inf) == (0, True)
        assert lib.verify_kzg_proof(inf, Z[:32], (1).to_bytes(32, "big"), inf) == (0, False)
    # duplicate proofs in one batch (same bucket collisions, P+P paths)
    Cd, Zd, Yd, PId = C[:48] * 6, Z[:32] * 6, Y[:32] * 6, PI[:48] * 6
    assert ctx.verify_kzg_proof_batch(Cd, Zd, Yd, PId, 6) == oracle.verify_kzg_proof_batch(Cd, Zd, Yd, PId, 6) == (0, True)
    assert ctx.last_artifacts()["A"] == oracle.last_artifacts()["A"]


def check_degenerate(gpu_ctx, oracle_ctx, n=40):
    """Infinity points, repeated proofs, zero evaluations: exceptional group-law paths on both sides."""
    inf = b.g1_compress(None)
    C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A4720, 0, n)
    # (1) every proof identical (all bucket additions are doublings / same-point adds)
    Cd, Zd, Yd, PId = C[:48] * n, Z[:32] * n, Y[:32] * n, PI[:48] * n
    assert gpu_ctx.verify_kzg_proof_batch(Cd, Zd, Yd, PId, n) == oracle_ctx.verify_kzg_proof_batch(Cd, Zd, Yd, PId, n) == (0, True)
    assert gpu_ctx.last_artifacts()["A"] == oracle_ctx.last_artifacts()["A"]
    # (2) all-infinity commitments and proofs with y = 0 (valid: the zero polynomial)
    Ci, PIi, Yi = inf * n, inf * n, bytes(32) * n
    assert gpu_ctx.verify_kzg_proof_batch(Ci, Z, Yi, PIi, n) == oracle_ctx.verify_kzg_proof_batch(Ci, Z, Yi, PIi, n) == (0, True)
    a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
    assert a1["A"] == a2["A"] == bytes(96) and a1["B"] == a2["B"] == bytes(96)
    # (3) same but y != 0 somewhere: A = -(r_j y_j) G1 != O -> reject
    jj = min(7, n - 1)
    Yj = bytes(32) * jj + (5).to_bytes(32, "big") + bytes(32) * (n - 1 - jj)
    assert gpu_ctx.verify_kzg_proof_batch(Ci, Z, Yj, PIi, n) == oracle_ctx.verify_kzg_proof_batch(Ci, Z, Yj, PIi, n) == (0, False)
    assert gpu_ctx.last_artifacts()["A"] == oracle_ctx.last_artifacts()["A"]
    # (4) a mix of infinity and regular proofs
    Cm = inf + C[48:]
    PIm = inf + PI[48:]
    Ym = bytes(32) + Y[32:]
    assert gpu_ctx.verify_kzg_proof_batch(Cm, Z, Ym, PIm, n) == oracle_ctx.verify_kzg_proof_batch(Cm, Z, Ym, PIm, n) == (0, True)
    assert gpu_ctx.last_artifacts()["A"] == oracle_ctx.last_artifacts()["A"]
    # (5) P and -P in the same batch positions (commitment negated -> wrong proof, still well-formed)
    negC0 = bytes([C[0] ^ 0x20]) + C[1:48]
    Cn = negC0 + C[48:]
    r1 = gpu_ctx.verify_kzg_proof_batch(Cn, Z, Y, PI, n)
    assert r1 == oracle_ctx.verify_kzg_proof_batch(Cn, Z, Y, PI, n) == (0, False)
    assert gpu_ctx.last_artifacts()["A"] == oracle_ctx.last_artifacts()["A"]


def check_status_classes(gpu_ctx, oracle_ctx, n=64):
    """Each malformed class planted into an otherwise valid batch gives BADARGS with the same counts."""
    rnd = random.Random(5)
    C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A4721, 0, n)
    for enc, st in negative_g1_encodings(rnd):
        if st == 0:
            continue
        j = rnd.randrange(n)
        for which in ("C", "PI"):
            Cb = C[:48 * j] + enc + C[48 * j + 48:] if which == "C" else C
            Pb = PI[:48 * j] + enc + PI[48 * j + 48:] if which == "PI" else PI
            assert gpu_ctx.verify_kzg_proof_batch(Cb, Z, Y, Pb, n) == oracle_ctx.verify_kzg_proof_batch(Cb, Z, Y, Pb, n) == (1, False)
            a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
            assert a1["n_bad_points"] == a2["n_bad_points"] == 1 and a1["n_bad_scalars"] == a2["n_bad_scalars"] == 0


def small_order_points(rnd):
    """Points of E(Fp) of small prime order l | h1 = 3 * 11^2 * 10177^2 * 859267^2 * 52437899^2 (cofactor torsion)."""
    order = b.H1 * R
    out = {3: (0, 2)}
    for ell in (11, 10177, 859267, 52437899):           # l^2 | h1: clear everything but the l-Sylow subgroup
        while ell not in out:
            t = b.g1_mul(order // (ell * ell), rand_curve_point(rnd))
            if t is None:
                continue
            t2 = b.g1_mul(ell, t)
            if t2 is not None:                          # order l^2 (cyclic Sylow): take the order-l multiple
                t = t2
            assert b.g1_mul(ell, t) is None
            out[ell] = t
    return out


def check_subgroup_batch(ctx, oracle, n=24, min_batch=2, ells=(3, 11, 10177, 859267, 52437899)):
    """Batched subgroup check (slice sums of the S1 / S3 buckets) against the oracle's per-point check: points
    with a cofactor component of every small prime order, alone, in cancelling pairs and everywhere."""
    rnd = random.Random(77)
    assert ctx.set_subgroup_batch_min(min_batch) == 0
    try:
        tors = small_order_points(rnd)
        C, Z, Y, PI = oracle.synth_instance(0x4B5A4731, 0, n)
        assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == oracle.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
        a1, a2 = ctx.last_artifacts(), oracle.last_artifacts()
        for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
            assert a1[key] == a2[key], key

        def shifted(buf, j, t):
            st, p = b.g1_decompress(buf[48 * j:48 * j + 48])
            assert st == 0
            q = b.g1_add(p, t)
            return buf[:48 * j] + b.g1_compress(q) + buf[48 * j + 48:]

        def expect_bad(Cb, Pb, count):
            assert ctx.verify_kzg_proof_batch(Cb, Z, Y, Pb, n) == oracle.verify_kzg_proof_batch(Cb, Z, Y, Pb, n) == (1, False)
            b1, b2 = ctx.last_artifacts(), oracle.last_artifacts()
            assert b1["n_bad_points"] == b2["n_bad_points"] == count

        for ell in ells:
            t = tors[ell]
            j = rnd.randrange(n)
            expect_bad(shifted(C, j, t), PI, 1)                       # one commitment with an order-l component
            expect_bad(C, shifted(PI, j, t), 1)                       # one proof
            expect_bad(C[:48 * j] + b.g1_compress(t) + C[48 * j + 48:], PI, 1)   # the torsion point itself
            i2 = (j + 1 + rnd.randrange(n - 1)) % n
            expect_bad(shifted(shifted(C, j, t), i2, b.g1_neg(t)), PI, 2)        # components cancel in a plain sum
            expect_bad(shifted(C, j, t), shifted(PI, j, b.g1_neg(t)), 2)
        if n <= 64:
            t3 = tors[3]
            Call, Pall = C, PI
            for j in range(n):
                Call, Pall = shifted(Call, j, t3), shifted(Pall, j, t3)
            expect_bad(Call, Pall, 2 * n)
        expect_bad(shifted(C, 0, rand_curve_point(rnd)), PI, 1)       # generic off-subgroup point
        # a wrong proof inside G1 is still only a rejection, not malformed input
        Pw = PI[48:96] + PI[:48] + PI[96:]
        assert ctx.verify_kzg_proof_batch(C, Z, Y, Pw, n) == oracle.verify_kzg_proof_batch(C, Z, Y, Pw, n) == (0, False)
        # malformed encodings (flags, x >= p, off curve) and scalars >= r on the batched path; mixed with an
        # off-subgroup point the counts still agree
        if n <= 4096:
            check_status_classes(ctx, oracle, n=n)
        for enc, st in negative_g1_encodings(rnd):
            if st in (1, 2, 3):
                j = rnd.randrange(1, n)
                expect_bad(shifted(C, 0, tors[3])[:48 * j] + enc + C[48 * j + 48:], PI, 2)
                break
        Zb = Z[:32] + b"\xff" * 32 + Z[64:]
        assert ctx.verify_kzg_proof_batch(C, Zb, Y, PI, n) == oracle.verify_kzg_proof_batch(C, Zb, Y, PI, n) == (1, False)
        assert ctx.last_artifacts()["n_bad_scalars"] == oracle.last_artifacts()["n_bad_scalars"] == 1
    finally:
        ctx.set_subgroup_batch_min(2)


def check_cell_batch(ctx, oracle, inst):
    """inst = (commitments, commitment_indices, cell_indices, cells, proofs) from the oracle generator."""
    comms, ci, xi, cells, proofs = inst
    m = len(ci)
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == oracle.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    a1, a2 = ctx.last_artifacts(), oracle.last_artifacts()
    assert a1["A"] == a2["A"] and a1["B"] == a2["B"] and a1["root"] == a2["root"]
    bad = bytearray(cells); bad[2048 * (m - 1) + 32 * 7 + 31] ^= 1                 # one evaluation tampered
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == oracle.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == (0, False)
    a1, a2 = ctx.last_artifacts(), oracle.last_artifacts()
    assert a1["A"] == a2["A"] and a1["B"] == a2["B"]
    xi2 = list(xi); xi2[0] = (xi2[0] + 5) % 128                                      # wrong coset
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi2, cells, proofs) == oracle.verify_cell_kzg_proof_batch(comms, ci, xi2, cells, proofs) == (0, False)
    if m >= 2:
        pr2 = proofs[48:96] + proofs[:48] + proofs[96:]
        assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, pr2) == oracle.verify_cell_kzg_proof_batch(comms, ci, xi, cells, pr2) == (0, False)
    # malformed inputs
    bad = bytearray(cells); bad[32:64] = b"\xff" * 32
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == oracle.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == (1, False)
    xi3 = list(xi); xi3[-1] = 128
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi3, cells, proofs) == (1, False)
    ci3 = list(ci); ci3[-1] = len(comms) // 48
    assert ctx.verify_cell_kzg_proof_batch(comms, ci3, xi, cells, proofs) == (1, False)
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, b.g1_compress((0, 2)) + proofs[48:]) == (1, False)
    assert ctx.verify_cell_kzg_proof_batch(b.g1_compress((0, 2)) + comms[48:], ci, xi, cells, proofs) == (1, False)


def check_random_differential(gpu_ctx, oracle_ctx, sizes=(1, 2, 3, 17, 64, 127, 128, 129, 255, 256, 1000, 1024, 1025, 2049, 3000), trials=40, pool=3000):
    """Randomised differential test: random batch sizes and random single-byte / structural corruptions;
    return code, verdict and (when well-formed) the pairing inputs must equal the oracle's."""
    rnd = random.Random(20261018)
    C0, Z0, Y0, PI0 = oracle_ctx.synth_instance(0x4B5A4730, 0, pool)
    specials = [b.g1_compress(None), b.g1_compress((0, 2)), bytes(48), b"\xff" * 48, b.g1_compress(b.G1), b.g1_compress(b.g1_neg(b.G1))]
    outcomes = {}
    for trial in range(trials):
        n = rnd.choice(list(sizes))
        off = rnd.randrange(0, pool - n + 1)
        arrs = [bytearray(C0[48 * off:48 * (off + n)]), bytearray(Z0[32 * off:32 * (off + n)]), bytearray(Y0[32 * off:32 * (off + n)]),
                bytearray(PI0[48 * off:48 * (off + n)])]
        widths = [48, 32, 32, 48]
        kind = rnd.choice(["none", "flip", "flip", "special", "swap", "dup", "highbit"])
        if kind == "flip":
            a = rnd.randrange(4)
            arrs[a][rnd.randrange(len(arrs[a]))] ^= 1 << rnd.randrange(8)
        elif kind == "special":
            a = rnd.choice([0, 3])
            j = rnd.randrange(n)
            arrs[a][48 * j:48 * j + 48] = rnd.choice(specials)
        elif kind == "swap" and n >= 2:
            a = rnd.randrange(4)
            w = widths[a]
            i, j = rnd.sample(range(n), 2)
            arrs[a][w * i:w * i + w], arrs[a][w * j:w * j + w] = arrs[a][w * j:w * j + w], arrs[a][w * i:w * i + w]
        elif kind == "dup" and n >= 2:
            i, j = rnd.sample(range(n), 2)
            for a, w in enumerate(widths):
                arrs[a][w * j:w * j + w] = arrs[a][w * i:w * i + w]
        elif kind == "highbit":
            a = rnd.choice([1, 2])
            j = rnd.randrange(n)
            arrs[a][32 * j] |= 0x80                       # scalar >= r
        args = [bytes(x) for x in arrs] + [n]
        r1 = gpu_ctx.verify_kzg_proof_batch(*args)
        r2 = oracle_ctx.verify_kzg_proof_batch(*args)
        assert r1 == r2, (trial, kind, n, r1, r2)
        outcomes[r1] = outcomes.get(r1, 0) + 1
        a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
        assert a1["n_bad_points"] == a2["n_bad_points"] and a1["n_bad_scalars"] == a2["n_bad_scalars"], (trial, kind)
        if r1[0] == 0:
            for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
                assert a1[key] == a2[key], (trial, kind, key)
    assert len(outcomes) >= 2, outcomes          # at least two of accepted / rejected / malformed were exercised


# ---- blob batch (include/kzgb200.h "Blob batch")
BLOB_LEN = 4096
OMEGA_4096 = pow(7, (R - 1) // 4096, R)


def _brp12(i):
    return int(format(i, "012b")[::-1], 2)


def blob_domain_point(i):
    return pow(OMEGA_4096, _brp12(i), R)


def poly_eval(coeffs, x):
    """coeffs: {degree: coefficient}"""
    return sum(a * pow(x, d, R) for d, a in coeffs.items()) % R


def python_blob_challenge(blob, comm):
    leaves = b"".join(hashlib.sha256(b"KZGB200/bleaf_v1" + blob[1024 * j:1024 * j + 1024]).digest() for j in range(128))
    return int.from_bytes(hashlib.sha256(b"KZGB200/blobz_v1" + comm + leaves).digest(), "big") % R


def python_blob(rnd):
    """One blob from a sparse polynomial of degree 4095, with commitment and proof from the known test tau."""
    coeffs = {d: rnd.randrange(R) for d in (0, 1, 2, 77, 2048, 4095)}
    blob = b"".join(poly_eval(coeffs, blob_domain_point(i)).to_bytes(32, "big") for i in range(BLOB_LEN))
    pt = poly_eval(coeffs, k.TAU)
    comm = b.g1_compress(b.g1_mul(pt, b.G1))
    z = python_blob_challenge(blob, comm)
    y = poly_eval(coeffs, z)
    proof = b.g1_compress(b.g1_mul((pt - y) * pow(k.TAU - z, R - 2, R) % R, b.G1))
    return blob, comm, proof, coeffs


def synth_blobs(oracle_lib, seed, m):
    blobs, comms, proofs = C.create_string_buffer(131072 * m), C.create_string_buffer(48 * m), C.create_string_buffer(48 * m)
    f = oracle_lib.lib.kzgb_oracle_synth_blobs
    f.argtypes, f.restype = [C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int], C.c_int
    assert f(seed, m, blobs, comms, proofs, 0) == 0
    return blobs.raw, comms.raw, proofs.raw


def check_blob_batch(ctx, oracle, inst):
    blobs, comms, proofs = inst
    m = len(comms) // 48
    rc1, z1, y1 = ctx.blob_challenges_evals(blobs, comms)
    rc2, z2, y2 = oracle.blob_challenges_evals(blobs, comms)
    assert rc1 == rc2 == 0 and z1 == z2 and y1 == y2
    assert ctx.verify_blob_kzg_proof_batch(blobs, comms, proofs) == oracle.verify_blob_kzg_proof_batch(blobs, comms, proofs) == (0, True)
    a1, a2 = ctx.last_artifacts(), oracle.last_artifacts()
    assert a1["A"] == a2["A"] and a1["B"] == a2["B"] and a1["root"] == a2["root"]
    # the blob path ends in the plain batch on (C, z, y, proofs)
    assert ctx.verify_kzg_proof_batch(comms, z1, y1, proofs, m) == (0, True)
    bad = bytearray(blobs); bad[131072 * (m - 1) + 32 * 1000 + 31] ^= 1            # one evaluation changed
    assert ctx.verify_blob_kzg_proof_batch(bytes(bad), comms, proofs) == oracle.verify_blob_kzg_proof_batch(bytes(bad), comms, proofs) == (0, False)
    if m >= 2:
        pr2 = proofs[48:96] + proofs[:48] + proofs[96:]
        assert ctx.verify_blob_kzg_proof_batch(blobs, comms, pr2) == oracle.verify_blob_kzg_proof_batch(blobs, comms, pr2) == (0, False)
    bad = bytearray(blobs); bad[32 * 5:32 * 6] = (R).to_bytes(32, "big")                # element == r
    assert ctx.verify_blob_kzg_proof_batch(bytes(bad), comms, proofs) == oracle.verify_blob_kzg_proof_batch(bytes(bad), comms, proofs) == (1, False)
    assert ctx.verify_blob_kzg_proof_batch(blobs, b.g1_compress((0, 2)) + comms[48:], proofs) == (1, False)   # off-subgroup commitment
    # caller-chosen evaluation points, including points of the domain
    pts = [blob_domain_point(3), blob_domain_point(4095), 0, 1, R - 1][:max(1, min(5, m))]
    zin = b"".join(v.to_bytes(32, "big") for v in pts)
    r1, r2 = ctx.blob_eval(blobs[:131072 * len(pts)], zin), oracle.blob_eval(blobs[:131072 * len(pts)], zin)
    assert r1 == r2 and r1[0] == 0
    assert r1[1][:32] == blobs[32 * 3:32 * 4]



def check_pipeline(ctx, oracle, depth=3, sizes=(700, 64, 3000, 1, 129, 2048, 5, 1500), seed=0x4B5A4761):
    """verify_kzg_proof_batch_submit / _wait: interleaved batches of different sizes, some with a planted wrong proof or a
    malformed element, `depth` in flight; every ticket returns what the oracle's plain call returns; ticket rules."""
    assert ctx.pipeline_init(depth) == 0
    jobs = []
    for k, n in enumerate(sizes):
        C, Z, Y, PI = oracle.synth_instance(seed + k, 0, n)
        if k % 3 == 1 and n >= 2:
            PI = PI[48:96] + PI[:48] + PI[96:]                      # two proofs swapped: well-formed, wrong
        if k % 4 == 2:
            Z = Z[:32 * (n // 2)] + b"\xff" * 32 + Z[32 * (n // 2 + 1):]      # scalar >= r: malformed
        jobs.append((C, Z, Y, PI, n))
    want = [oracle.verify_kzg_proof_batch(*j) for j in jobs]
    assert {w for w in want} >= {(0, True), (0, False), (1, False)}
    got, tickets = [None] * len(jobs), []
    for k, j in enumerate(jobs):
        if len(tickets) == depth:                                   # the pipeline is full: a further submit is refused
            assert ctx.verify_kzg_proof_batch_submit(*j)[0] == 1
            t, kk = tickets.pop(0)
            got[kk] = ctx.verify_kzg_proof_batch_wait(t)
            assert ctx.verify_kzg_proof_batch_wait(t) == (1, False)     # a ticket is collected once
        rc, t = ctx.verify_kzg_proof_batch_submit(*j)
        assert rc == 0
        tickets.append((t, k))
    for t, kk in reversed(tickets):                                 # the rest in reverse order
        got[kk] = ctx.verify_kzg_proof_batch_wait(t)
    assert got == want, (got, want)
    assert ctx.verify_kzg_proof_batch_wait(10 ** 6) == (1, False)   # unknown ticket
    # the plain entry point still works beside the pipeline
    assert ctx.verify_kzg_proof_batch(*jobs[0]) == want[0]
    assert ctx.pipeline_init(1) == 0
