"""CPU oracle (C++) vs the pure-Python model: pins the oracle before anything is compared to it."""
import hashlib
import random
import struct

import pytest

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k
from tests.helpers import (P, R, edge_fps, f12_bytes, f12_from_bytes, fp_bytes, fr_bytes, negative_g1_encodings,
                           rand_curve_point, rand_g1)


def test_exports(oracle_lib):
    for name in oracle_lib.EXPORTS:
        assert hasattr(oracle_lib.lib, name), name
    assert "oracle" in oracle_lib.version()


def test_fp_fr_ops(oracle_ctx):
    rnd = random.Random(11)
    vals = edge_fps() + [rnd.randrange(P) for _ in range(40)]
    pairs = [(x, y) for x in vals[:13] for y in vals[:13]] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(200)]
    data = b"".join(fp_bytes(x) + fp_bytes(y) for x, y in pairs)
    for op, fn in (("FP_MUL", lambda x, y: x * y), ("FP_ADD", lambda x, y: x + y), ("FP_SUB", lambda x, y: x - y)):
        rc, out = oracle_ctx.debug_op(op, data)
        assert rc == 0
        assert out == b"".join(fp_bytes(fn(x, y)) for x, y in pairs), op
    data = b"".join(fp_bytes(x) for x in vals)
    rc, out = oracle_ctx.debug_op("FP_SQR", data)
    assert out == b"".join(fp_bytes(x * x) for x in vals)
    rc, out = oracle_ctx.debug_op("FP_INV", data)
    assert out == b"".join(fp_bytes(pow(x, P - 2, P)) for x in vals)
    rc, out = oracle_ctx.debug_op("FP_SQRT_CAND", data)
    assert out == b"".join(fp_bytes(pow(x, (P + 1) // 4, P)) for x in vals)
    frs = [0, 1, R - 1, 2 ** 128 - 1, 2 ** 254] + [rnd.randrange(R) for _ in range(100)]
    prs = [(x, y) for x in frs[:5] for y in frs[:5]] + [(rnd.choice(frs), rnd.randrange(R)) for _ in range(100)]
    data = b"".join(fr_bytes(x) + fr_bytes(y) for x, y in prs)
    rc, out = oracle_ctx.debug_op("FR_MUL", data)
    assert out == b"".join(fr_bytes(x * y) for x, y in prs)
    rc, out = oracle_ctx.debug_op("FR_ADD", data)
    assert out == b"".join(fr_bytes(x + y) for x, y in prs)
    # non-canonical operand is rejected
    rc, _ = oracle_ctx.debug_op("FP_SQR", P.to_bytes(48, "big"))
    assert rc == 1


def test_sha256(oracle_ctx):
    rnd = random.Random(3)
    msgs = [bytes(rnd.randrange(256) for _ in range(64)) for _ in range(8)] + [bytes(64), b"\xff" * 64]
    rc, out = oracle_ctx.debug_op("SHA256_64", b"".join(msgs))
    assert out == b"".join(hashlib.sha256(m).digest() for m in msgs)


def test_g1_ops(oracle_ctx):
    rnd = random.Random(5)
    pts = [rand_g1(rnd) for _ in range(6)] + [None, rand_curve_point(rnd), (0, 2)]
    pairs = [(p, q) for p in pts for q in pts] + [(pts[0], b.g1_neg(pts[0]))]
    data = b"".join(b.g1_affine_bytes(p) + b.g1_affine_bytes(q) for p, q in pairs)
    rc, out = oracle_ctx.debug_op("G1_ADD", data)
    assert rc == 0 and out == b"".join(b.g1_affine_bytes(b.g1_add(p, q)) for p, q in pairs)
    data = b"".join(b.g1_affine_bytes(p) for p in pts)
    rc, out = oracle_ctx.debug_op("G1_DBL", data)
    assert out == b"".join(b.g1_affine_bytes(b.g1_add(p, p)) for p in pts)
    rc, out = oracle_ctx.debug_op("G1_MUL_XSQ", data)
    assert out == b"".join(b.g1_affine_bytes(b.g1_mul(b.X * b.X, p)) for p in pts)
    ks = [0, 1, 2, R - 1, 2 ** 128 - 1] + [rnd.randrange(R) for _ in range(4)]
    data = b"".join(b.g1_affine_bytes(pts[i % 6]) + fr_bytes(kk) for i, kk in enumerate(ks))
    rc, out = oracle_ctx.debug_op("G1_MUL", data)
    assert out == b"".join(b.g1_affine_bytes(b.g1_mul(kk, pts[i % 6])) for i, kk in enumerate(ks))


def test_decompress_statuses(oracle_lib, oracle_ctx):
    rnd = random.Random(9)
    cases = negative_g1_encodings(rnd) + [(b.g1_compress(rand_g1(rnd)), 0) for _ in range(5)]
    data = b"".join(c for c, _ in cases)
    rc, aff, st = oracle_ctx.g1_decompress_batch(data)
    assert rc == 0
    assert list(st) == [s for _, s in cases]
    for i, (enc, s) in enumerate(cases):
        ms, mp = b.g1_decompress(enc)
        assert ms == s
        assert aff[96 * i:96 * i + 96] == b.g1_affine_bytes(mp)
        assert oracle_lib.lib.kzgb_oracle_g1_status_slow(enc) == s       # fast test == slow [r]P test


def test_tower_and_pairing(oracle_ctx):
    rnd = random.Random(21)
    a = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(6)]
    c = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(6)]
    rc, out = oracle_ctx.debug_op("FP12_MUL", f12_bytes(a) + f12_bytes(c))
    assert f12_from_bytes(out) == b.f12_mul(a, c)
    rc, out = oracle_ctx.debug_op("FP12_FROB1", f12_bytes(a))
    assert f12_from_bytes(out) == b.f12_frob(a, 1)
    rc, out = oracle_ctx.debug_op("FP12_FROB2", f12_bytes(a))
    assert f12_from_bytes(out) == b.f12_frob(a, 2)
    rc, out = oracle_ctx.debug_op("FP12_INV", f12_bytes(a))
    assert f12_from_bytes(out) == b.f12_inv(a)
    # final exponentiation: oracle computes the cube of the textbook exponent (HHT hard part)
    rc, out = oracle_ctx.debug_op("FINAL_EXP", f12_bytes(a))
    fe = b.final_exp(a)
    assert f12_from_bytes(out) == b.f12_mul(b.f12_mul(fe, fe), fe)
    # full two-pairing value against the model (ctx setup = test tau)
    g2t = b.g2_mul(k.TAU, b.G2)
    A, B = rand_g1(rnd), rand_g1(rnd)
    rc, out = oracle_ctx.debug_op("MILLER_FE", b.g1_affine_bytes(A) + b.g1_affine_bytes(B))
    want = b.final_exp(b.f12_mul(b.miller_loop(A, b.G2), b.miller_loop(B, g2t)))
    assert f12_from_bytes(out) == b.f12_mul(b.f12_mul(want, want), want)
    # verdicts: A + tau B = O accepted, anything else rejected, infinity pair accepted
    Bp = rand_g1(rnd)
    Ap = b.g1_neg(b.g1_mul(k.TAU, Bp))
    assert oracle_ctx.pairing_check(b.g1_affine_bytes(Ap), b.g1_affine_bytes(Bp)) == (0, True)
    assert oracle_ctx.pairing_check(b.g1_affine_bytes(A), b.g1_affine_bytes(B)) == (0, False)
    assert oracle_ctx.pairing_check(bytes(96), bytes(96)) == (0, True)
    assert oracle_ctx.pairing_check(b.g1_affine_bytes(A), bytes(96)) == (0, False)


def test_setup_and_generator_match_model(oracle_lib, oracle_ctx):
    g1, g2 = oracle_lib.synth_setup(3, 3)
    assert g1 == k.setup_g1(3) and g2 == k.setup_g2(3)
    seed = 0x4B5A4701
    C, Z, Y, PI = oracle_ctx.synth_instance(seed, 0, 6)
    assert (C, Z, Y, PI) == k.gen_instance_shortcut(seed, 6)
    C2, Z2, Y2, PI2 = oracle_ctx.synth_instance(seed, 4, 2)
    assert C2 == C[4 * 48:] and PI2 == PI[4 * 48:] and Z2 == Z[4 * 32:]
    import ctypes
    n, nc = 2, 16
    bufs = [ctypes.create_string_buffer(s * n) for s in (48, 32, 32, 48)]
    oracle_lib.lib.kzgb_oracle_synth_instance_poly(ctypes.c_uint64(5), ctypes.c_size_t(n), ctypes.c_size_t(nc), *bufs, 1)
    assert tuple(x.raw for x in bufs) == k.gen_instance_poly(5, n, nc)
    assert oracle_lib.lib.kzgb_oracle_plant_index(ctypes.c_uint64(seed), ctypes.c_uint64(6)) == k.plant_index(seed, 6)


def test_fs_and_batch_artifacts_match_model(oracle_lib, oracle_ctx):
    seed, n = 0x4B5A4701, 9
    C, Z, Y, PI = k.gen_instance_shortcut(seed, n)
    rc, root, r = oracle_ctx.fs_challenges(C, Z, Y, PI, n)
    mroot, mr = k.fs_challenges(C, Z, Y, PI, n)
    assert rc == 0 and root == mroot and r == b"".join(v.to_bytes(16, "big") for v in mr)
    rc, ok = oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n)
    assert (rc, ok) == (0, True)
    art = oracle_ctx.last_artifacts()
    m = k.batch_artifacts(C, Z, Y, PI, n)
    for key in ("S1", "S2", "S3", "A", "B"):
        assert art[key] == b.g1_affine_bytes(m[key]), key
    assert art["sum_ry"] == fr_bytes(m["sum_ry"]) and art["root"] == m["root"]
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(art["A"], art["B"]) == 1
    # planted invalid proof: well-formed, rejected
    j = k.plant_index(seed, n)
    PI_bad = k.plant_invalid(PI, j)
    assert oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI_bad, n) == (0, False)
    art = oracle_ctx.last_artifacts()
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(art["A"], art["B"]) == 0
    # single-proof entry point
    assert oracle_ctx.verify_kzg_proof(C[:48], Z[:32], Y[:32], PI[:48]) == (0, True)
    assert oracle_ctx.verify_kzg_proof(C[:48], Z[:32], Y[:32], PI[48:96]) == (0, False)
    # malformed inputs -> BADARGS
    assert oracle_ctx.verify_kzg_proof_batch(b.g1_compress((0, 2)) + C[48:], Z, Y, PI, n) == (1, False)
    assert oracle_ctx.verify_kzg_proof_batch(C, R.to_bytes(32, "big") + Z[32:], Y, PI, n) == (1, False)
    assert oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, 0) == (1, False)


def test_msm_matches_model(oracle_ctx):
    rnd = random.Random(17)
    for m, nbits in ((1, 255), (5, 255), (40, 255), (40, 128)):
        pts = [rand_g1(rnd) if rnd.random() < 0.9 else None for _ in range(m)]
        if m >= 5:
            pts[1] = pts[0]                       # duplicate point
            pts[2] = b.g1_neg(pts[0])             # and its negation
        ks = [rnd.randrange(2 ** 128 if nbits == 128 else R) for _ in range(m)]
        if m >= 5:
            ks[1] = ks[0]; ks[2] = ks[0]; ks[3] = 0; ks[4] = (2 ** 128 - 1) if nbits == 128 else R - 1
        want = None
        for p_, k_ in zip(pts, ks):
            want = b.g1_add(want, b.g1_mul(k_, p_))
        rc, out = oracle_ctx.g1_msm(b"".join(b.g1_affine_bytes(p_) for p_ in pts), b"".join(fr_bytes(k_) for k_ in ks), nbits)
        assert rc == 0 and out == b.g1_affine_bytes(want)


def test_sharded_equals_single(oracle_lib):
    """Shard-count invariance on the oracle: 1 vs 2 vs 3 shards, n not a multiple of 1024."""
    ctx = oracle_lib.test_context(devices=[0, 0, 0])
    seed, n = 0x4B5A4702, 2048 + 100
    C, Z, Y, PI = ctx.synth_instance(seed, 0, n)
    rc, ok = ctx.verify_kzg_proof_batch(C, Z, Y, PI, n)
    assert (rc, ok) == (0, True)
    ref = ctx.last_artifacts()
    for bounds in ([0, n], [0, 384, n], [0, 1024, 2048, n]):
        digs, parts = b"", b""
        for s in range(len(bounds) - 1):
            lo, hi = bounds[s], bounds[s + 1]
            rc, d, nbad = ctx.shard_phase1(s, C[48 * lo:48 * hi], Z[32 * lo:32 * hi], Y[32 * lo:32 * hi], PI[48 * lo:48 * hi], hi - lo)
            assert rc == 0 and nbad == 0
            digs += d
        root = ctx.fs_root(digs, n)
        assert root == ref["root"]
        for s in range(len(bounds) - 1):
            rc, p = ctx.shard_phase2(s, root, bounds[s])
            assert rc == 0
            parts += p
        assert ctx.combine_verify(parts) == (0, True)
        art = ctx.last_artifacts()
        assert art["A"] == ref["A"] and art["B"] == ref["B"] and art["sum_ry"] == ref["sum_ry"]
    ctx.close()


def test_config0_real_polynomials_n64(oracle_lib, oracle_ctx):
    """BASELINE.json config[0]: 64 synthetic proofs from real degree-4095 polynomials (known test tau) on host cores."""
    import ctypes
    n, ncoef, seed = 64, 4096, 0x4B5A4700
    bufs = [ctypes.create_string_buffer(s_ * n) for s_ in (48, 32, 32, 48)]
    assert oracle_lib.lib.kzgb_oracle_synth_instance_poly(ctypes.c_uint64(seed), ctypes.c_size_t(n), ctypes.c_size_t(ncoef), *bufs, 0) == 0
    C, Z, Y, PI = (x.raw for x in bufs)
    assert oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    art = oracle_ctx.last_artifacts()
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(art["A"], art["B"]) == 1
    # y_j tampered: f_j(z_j) + 1 is not the evaluation
    Ybad = Y[:32 * 5] + ((int.from_bytes(Y[32 * 5:32 * 6], "big") + 1) % R).to_bytes(32, "big") + Y[32 * 6:]
    assert oracle_ctx.verify_kzg_proof_batch(C, Z, Ybad, PI, n) == (0, False)
    for i in (0, 63):
        assert oracle_ctx.verify_kzg_proof(C[48 * i:48 * i + 48], Z[32 * i:32 * i + 32], Y[32 * i:32 * i + 32], PI[48 * i:48 * i + 48]) == (0, True)


def test_pipeline_ticket_rules(oracle_lib):
    """submit / wait of the ABI on the oracle (verified at submit, parked until collected): the ticket rules the GPU
    test relies on -- refusal when `depth` tickets are uncollected, single collection, any order."""
    from tests import parity_suite as ps
    ctx = oracle_lib.test_context()
    full = oracle_lib.test_context()
    assert ctx.verify_kzg_proof_batch_submit(b"\0" * 48, b"\0" * 32, b"\0" * 32, b"\0" * 48, 1)[0] == 1      # no pipeline yet
    ps.check_pipeline(ctx, full, depth=3, sizes=(70, 64, 130, 1, 129, 33, 5, 150))
    ctx.close(); full.close()
