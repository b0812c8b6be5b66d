"""CUDA library vs CPU oracle through the C ABI (include/kzgb200.h) -- the parity tests proper.
Run on a B200 with `pytest -m gpu`.  Bit-exact or fail (BASELINE.json:5)."""
import ctypes
import random

import pytest

from oracle.pymodel import bls12_381 as b
from tests import parity_suite as ps
from tests.helpers import rand_g1

pytestmark = pytest.mark.gpu


def test_gpu_lib_is_the_cuda_build(gpu_lib):
    assert "cuda" in gpu_lib.version()


def test_gpu_field_ops(gpu_ctx, oracle_ctx):
    ps.check_field_ops(gpu_ctx, oracle_ctx, n_random=2000)


def test_gpu_per_point_subgroup_path(gpu_lib, oracle_ctx, oracle_lib):
    """The deterministic per-point subgroup check (kzgb_set_subgroup_batch_min(0)) stays a first-class path: the
    batched check falls back to it, verify_kzg_proof and the cell commitments use it."""
    ctx = gpu_lib.test_context(n_max=1 << 16)
    assert ctx.set_subgroup_batch_min(0) == 0
    ps.check_verify(ctx, oracle_ctx, oracle_lib, sizes=(1, 2, 9, 1000))
    ps.check_status_classes(ctx, oracle_ctx, n=64)
    ps.check_degenerate(ctx, oracle_ctx, n=40)
    ctx.close()


def test_gpu_subgroup_batch_small(gpu_ctx, oracle_ctx):
    ps.check_subgroup_batch(gpu_ctx, oracle_ctx, n=24)


def test_gpu_subgroup_batch_default_threshold(gpu_ctx, oracle_ctx):
    """n = 2^15 proofs: the batched check is the default path; a single order-3 / order-11 component is found."""
    ps.check_subgroup_batch(gpu_ctx, oracle_ctx, n=1 << 15, min_batch=16384, ells=(3, 11))


def test_gpu_blob_batch(gpu_ctx, oracle_ctx, oracle_lib):
    """Blob-level caller: z, y bit-exact vs the oracle, verdicts, malformed blobs; 600 blobs
    fill both 256-blob staging buffers and reuse the first."""
    ps.check_blob_batch(gpu_ctx, oracle_ctx, ps.synth_blobs(oracle_lib, 0x4B5A4743, 600))
    import random
    blob, comm, proof, coeffs = ps.python_blob(random.Random(12))          # against direct polynomial evaluation
    rc, zs, ys = gpu_ctx.blob_challenges_evals(blob, comm)
    z = ps.python_blob_challenge(blob, comm)
    assert rc == 0 and zs == z.to_bytes(32, "big") and ys == ps.poly_eval(coeffs, z).to_bytes(32, "big")
    assert gpu_ctx.verify_blob_kzg_proof_batch(blob, comm, proof) == (0, True)


@pytest.mark.parametrize("n", [16383, 16384, 16385, 17001, 33333])
def test_gpu_sizes_around_the_batched_check_threshold(gpu_ctx, oracle_ctx, oracle_lib, n):
    """Ragged sizes on both sides of the default threshold of the batched subgroup check (16384): verdict and every
    artefact equal the oracle's; a planted wrong proof is rejected with equal pairing inputs."""
    seed = 0x4B5A4750 + n
    C, Z, Y, PI = oracle_ctx.synth_instance(seed, 0, n)
    assert gpu_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert a1[key] == a2[key], key
    j = n - 2
    PIb = PI[:48 * j] + PI[48 * (j + 1):48 * (j + 2)] + PI[48 * j:48 * (j + 1)] + PI[48 * (j + 2):]
    assert gpu_ctx.verify_kzg_proof_batch(C, Z, Y, PIb, n) == oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PIb, n) == (0, False)
    a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
    assert a1["A"] == a2["A"] and a1["B"] == a2["B"]


def test_gpu_product_has_no_fpd_ops(gpu_ctx):
    """The FP64-limb multiplier (recorded negative result) lives in tools/microbench, not in the product library."""
    assert gpu_ctx.debug_op("FPD_MUL", bytes(96))[0] == 1


def test_gpu_pipeline_submit_wait(gpu_lib, oracle_ctx):
    """Library-level pipelining (kzgb_pipeline_init / submit / wait): interleaved batches against the oracle."""
    ctx = gpu_lib.test_context(n_max=4096)
    ps.check_pipeline(ctx, oracle_ctx, depth=3)
    ps.check_pipeline(ctx, oracle_ctx, depth=2, sizes=(4096, 4096, 4095, 4096, 2, 4096), seed=0x4B5A4771)
    # device-resident inputs through the pipeline
    import torch
    n = 4096
    bufs = [[torch.empty(s * n, dtype=torch.uint8, device="cuda") for s in (48, 32, 32, 48)] for _ in range(4)]
    for k, b_ in enumerate(bufs):
        ctx.synth_instance(0x4B5A4781 + k, 0, n, device_ptrs=tuple(t.data_ptr() for t in b_))
    bufs[2][3][:48] = bufs[2][3][48:96]
    torch.cuda.synchronize()
    assert ctx.pipeline_init(2) == 0
    res, pend = [], []
    for b_ in bufs:
        if len(pend) == 2:
            res.append(ctx.verify_kzg_proof_batch_wait(pend.pop(0)))
        rc, t = ctx.verify_kzg_proof_batch_submit(*[t.data_ptr() for t in b_], n, on_device=True)
        assert rc == 0
        pend.append(t)
    res += [ctx.verify_kzg_proof_batch_wait(t) for t in pend]
    assert res == [(0, True), (0, True), (0, False), (0, True)]
    ctx.close()


def test_gpu_g1_ops(gpu_ctx, oracle_ctx):
    ps.check_g1_ops(gpu_ctx, oracle_ctx)


def test_gpu_decompress(gpu_ctx, oracle_ctx):
    ps.check_decompress(gpu_ctx, oracle_ctx, n_valid=300)


def test_gpu_tower_and_pairing(gpu_ctx, oracle_ctx):
    ps.check_tower_and_pairing(gpu_ctx, oracle_ctx)


def test_gpu_fs(gpu_ctx, oracle_ctx):
    ps.check_fs(gpu_ctx, oracle_ctx)


def test_gpu_msm(gpu_ctx, oracle_ctx):
    ps.check_msm(gpu_ctx, oracle_ctx)


def test_gpu_synth_matches_oracle_generator(gpu_ctx, oracle_ctx):
    ps.check_synth(gpu_ctx, oracle_ctx, n=300)


def test_gpu_verify_small(gpu_ctx, oracle_ctx, oracle_lib):
    ps.check_verify(gpu_ctx, oracle_ctx, oracle_lib, sizes=(1, 2, 9, 64, 1025))


def test_gpu_msm_4096_vs_oracle(gpu_ctx, oracle_ctx):
    """MSM over 4096 distinct subgroup points (from the generator), 255- and 128-bit scalars."""
    C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A4701, 0, 2048)
    rc, aff, st = oracle_ctx.g1_decompress_batch(C + PI)
    assert rc == 0 and not any(st)
    rnd = random.Random(77)
    for nbits in (255, 128):
        ks = b"".join(rnd.randrange(2 ** 128 if nbits == 128 else b.R).to_bytes(32, "big") for _ in range(4096))
        r1 = gpu_ctx.g1_msm(aff, ks, nbits)
        r2 = oracle_ctx.g1_msm(aff, ks, nbits)
        assert r1[0] == 0 and r1 == r2


def test_gpu_config_n4096(gpu_ctx, oracle_ctx):
    """BASELINE.json config[1]: n=4096 compressed inputs incl. decompression + subgroup checks."""
    n, seed = 4096, 0x4B5A4701
    C, Z, Y, PI = oracle_ctx.synth_instance(seed, 0, n)
    assert gpu_ctx.synth_instance(seed, 0, n) == (C, Z, Y, PI)
    assert gpu_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    assert oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert a1[key] == a2[key], key
    rc1, aff1, st1 = gpu_ctx.g1_decompress_batch(C + PI)
    rc2, aff2, st2 = oracle_ctx.g1_decompress_batch(C + PI)
    assert (rc1, aff1, st1) == (rc2, aff2, st2)


def test_gpu_config_n65536_planted_invalid(gpu_lib, oracle_lib):
    """BASELINE.json config[2]: n=2^16 with one planted invalid proof: must reject; valid batch accepts.
    Verdicts are cross-checked pairing-free with the known test tau (A + tau*B == O)."""
    n, seed = 1 << 16, 0x4B5A4702
    ctx = gpu_lib.test_context(n_max=n)
    octx = oracle_lib.test_context()
    C, Z, Y, PI = ctx.synth_instance(seed, 0, n)
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    art = ctx.last_artifacts()
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(art["A"], art["B"]) == 1
    assert octx.pairing_check(art["A"], art["B"]) == (0, True)
    # byte-exact diff of every artefact against the oracle at the size the metric is quoted on (window width c = 13 here)
    assert octx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    oart = octx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert art[key] == oart[key], key
    # spot parity of the oracle generator on a slice of the same stream
    assert octx.synth_instance(seed, 12345, 64) == tuple(x[s * 12345:s * (12345 + 64)] for x, s in ((C, 48), (Z, 32), (Y, 32), (PI, 48)))
    j = oracle_lib.lib.kzgb_oracle_plant_index(ctypes.c_uint64(seed), ctypes.c_uint64(n))
    bad = bytearray(PI)
    assert oracle_lib.lib.kzgb_oracle_plant_invalid((ctypes.c_uint8 * len(bad)).from_buffer(bad), ctypes.c_size_t(j)) == 0
    assert ctx.verify_kzg_proof_batch(C, Z, Y, bytes(bad), n) == (0, False)
    art = ctx.last_artifacts()
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(art["A"], art["B"]) == 0
    # the planted batch against the oracle: same verdict, same sums and pairing inputs, and the oracle's pairing on the
    # CUDA library's A, B rejects as well
    assert octx.verify_kzg_proof_batch(C, Z, Y, bytes(bad), n) == (0, False)
    oart = octx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert art[key] == oart[key], key
    assert octx.pairing_check(art["A"], art["B"]) == (0, False)
    # an off-subgroup point anywhere in the batch is BADARGS
    bad2 = bytearray(C)
    bad2[48 * 777:48 * 778] = b.g1_compress((0, 2))
    assert ctx.verify_kzg_proof_batch(bytes(bad2), Z, Y, PI, n) == (1, False)
    assert ctx.last_artifacts()["n_bad_points"] == 1
    ctx.close()
    octx.close()


def test_gpu_shards_equal_single_and_cross_library_combine(gpu_lib, oracle_lib):
    """Shard-count invariance on ONE device (G virtual shards run sequentially) and the
    cross-library check: partials produced by the CUDA library are accepted by the oracle's combine."""
    n, seed = 5 * 1024 + 100, 0x4B5A4703
    ctx = gpu_lib.test_context(devices=[0, 0, 0], n_max=n)
    octx = oracle_lib.test_context()
    C, Z, Y, PI = ctx.synth_instance(seed, 0, n)
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)        # 3 slots on device 0
    ref_root = ctx.last_artifacts()["root"]
    for bounds in ([0, n], [0, 2048, n], [0, 384, 4096, n]):
        digs = b""
        for s in range(len(bounds) - 1):
            lo, hi = bounds[s], bounds[s + 1]
            rc, d, _ = ctx.shard_phase1(s, C[48 * lo:48 * hi], Z[32 * lo:32 * hi], Y[32 * lo:32 * hi], PI[48 * lo:48 * hi], hi - lo)
            assert rc == 0
            digs += d
        root = ctx.fs_root(digs, n)
        assert root == ref_root == octx.fs_root(digs, n)
        parts = b""
        for s in range(len(bounds) - 1):
            rc, p = ctx.shard_phase2(s, root, bounds[s])
            assert rc == 0
            parts += p
        assert ctx.combine_verify(parts) == (0, True)
        assert octx.combine_verify(parts) == (0, True)
        a1, a2 = ctx.last_artifacts(), octx.last_artifacts()
        assert a1["A"] == a2["A"] and a1["B"] == a2["B"] and a1["sum_ry"] == a2["sum_ry"]
    ctx.close()
    octx.close()


def test_gpu_device_resident_entry_point(gpu_lib):
    import torch
    n, seed = 4096, 0x4B5A4704
    ctx = gpu_lib.test_context(n_max=n)
    bufs = [torch.empty(s * n, dtype=torch.uint8, device="cuda") for s in (48, 32, 32, 48)]
    ctx.synth_instance(seed, 0, n, device_ptrs=tuple(t.data_ptr() for t in bufs))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    assert ctx.verify_kzg_proof_batch_device(*[t.data_ptr() for t in bufs], n, stream) == (0, True)
    host = [bytes(t.cpu().numpy().tobytes()) for t in bufs]
    assert ctx.verify_kzg_proof_batch(*host, n) == (0, True)
    bufs[3][48 * 5 + 47] ^= 1           # corrupt one proof byte on the device: not a valid encoding any more (w.h.p.) or wrong proof
    rc, ok = ctx.verify_kzg_proof_batch_device(*[t.data_ptr() for t in bufs], n, stream)
    assert not ok
    assert ctx.launch_count() > 0
    ctx.close()


@pytest.mark.parametrize("n", [3, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025, 2047, 3000])
def test_gpu_verify_ragged_sizes(gpu_ctx, oracle_ctx, n):
    """Ragged batch sizes around the window / chunk boundaries: verdict + all artefacts bit-exact."""
    seed = 0x4B5A4710 + n
    C, Z, Y, PI = oracle_ctx.synth_instance(seed, 0, n)
    assert gpu_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert a1[key] == a2[key], (n, key)
    # swap two proofs: well-formed, must be rejected by both
    if n >= 2:
        PI2 = PI[48:96] + PI[:48] + PI[96:]
        assert gpu_ctx.verify_kzg_proof_batch(C, Z, Y, PI2, n) == oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI2, n) == (0, False)


def test_gpu_degenerate_batches(gpu_ctx, oracle_ctx):
    ps.check_degenerate(gpu_ctx, oracle_ctx)


def test_gpu_every_status_class_in_batch(gpu_ctx, oracle_ctx):
    ps.check_status_classes(gpu_ctx, oracle_ctx)


def test_gpu_msm_large_random_vs_linearity(gpu_lib):
    """Size-independent property at 2^16 points: MSM(k) + MSM(k') == MSM(k + k') (bit-exact affine bytes
    compared through a 2-point MSM with unit scalars)."""
    n = 1 << 16
    ctx = gpu_lib.test_context(n_max=n)
    C, Z, Y, PI = ctx.synth_instance(0x4B5A4722, 0, n)
    rc, aff, st = ctx.g1_decompress_batch(C)
    assert rc == 0 and not any(st)
    rnd = random.Random(9)
    half = (b.R - 1) // 2
    k1 = [rnd.randrange(half) for _ in range(n)]
    k2 = [rnd.randrange(half) for _ in range(n)]
    enc = lambda ks: b"".join(k.to_bytes(32, "big") for k in ks)
    rc1, s1 = ctx.g1_msm(aff, enc(k1), 255)
    rc2, s2 = ctx.g1_msm(aff, enc(k2), 255)
    rc3, s3 = ctx.g1_msm(aff, enc([x + y for x, y in zip(k1, k2)]), 255)
    assert rc1 == rc2 == rc3 == 0
    one = (1).to_bytes(32, "big")
    rc4, s12 = ctx.g1_msm(s1 + s2, one + one, 255)
    assert rc4 == 0 and s12 == s3 and s3 != bytes(96)
    ctx.close()


def test_gpu_config0_real_polynomials_n64(gpu_ctx, oracle_ctx, oracle_lib):
    """BASELINE.json config[0] inputs (64 proofs from real degree-4095 polynomials, generated by the oracle)
    through the CUDA library: same verdicts and artefacts."""
    n, ncoef, seed = 64, 4096, 0x4B5A4700
    bufs = [ctypes.create_string_buffer(s_ * n) for s_ in (48, 32, 32, 48)]
    assert oracle_lib.lib.kzgb_oracle_synth_instance_poly(ctypes.c_uint64(seed), ctypes.c_size_t(n), ctypes.c_size_t(ncoef), *bufs, 0) == 0
    C, Z, Y, PI = (x.raw for x in bufs)
    assert gpu_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    a1, a2 = gpu_ctx.last_artifacts(), oracle_ctx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert a1[key] == a2[key], key
    Ybad = Y[:32 * 5] + ((int.from_bytes(Y[32 * 5:32 * 6], "big") + 1) % b.R).to_bytes(32, "big") + Y[32 * 6:]
    assert gpu_ctx.verify_kzg_proof_batch(C, Z, Ybad, PI, n) == oracle_ctx.verify_kzg_proof_batch(C, Z, Ybad, PI, n) == (0, False)


def test_gpu_cell_batch_small(gpu_lib, oracle_lib):
    from tests.test_cells_oracle import synth_cells
    g1, g2 = oracle_lib.synth_setup(64, 65)
    ctx, octx = gpu_lib.context(g1, g2, n_max=4096), oracle_lib.context(g1, g2)
    ps.check_cell_batch(ctx, octx, synth_cells(oracle_lib, 0x4B5A4724, 2, 3, 200))
    ps.check_cell_batch(ctx, octx, synth_cells(oracle_lib, 0x4B5A4725, 5, 40, 4096))
    # the plain batch entry points keep working on an extended-setup context
    C, Z, Y, PI = octx.synth_instance(0x4B5A4726, 0, 33)
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, 33) == octx.verify_kzg_proof_batch(C, Z, Y, PI, 33) == (0, True)
    # and a context without the extended setup refuses the cell batch
    plain = gpu_lib.test_context(n_max=4096)
    inst = synth_cells(oracle_lib, 1, 1, 1, 64)
    assert plain.verify_cell_kzg_proof_batch(*inst) == (1, False)
    plain.close(); ctx.close(); octx.close()


def test_gpu_config4_cell_batch_128x128(gpu_lib, oracle_lib):
    """BASELINE.json config[4]: 128 blobs x 128 cells = 2^14 multi-point openings (one GPU; the work is ~ms)."""
    import time
    from tests.test_cells_oracle import synth_cells
    g1, g2 = oracle_lib.synth_setup(64, 65)
    ctx, octx = gpu_lib.context(g1, g2, n_max=1 << 15), oracle_lib.context(g1, g2)
    comms, ci, xi, cells, proofs = synth_cells(oracle_lib, 0x4B5A4704, 128, 128, 4096)
    assert len(ci) == 1 << 14
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    t0 = time.perf_counter()
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    gpu_s = time.perf_counter() - t0
    a1 = ctx.last_artifacts()
    t0 = time.perf_counter()
    assert octx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    cpu_s = time.perf_counter() - t0
    a2 = octx.last_artifacts()
    assert a1["A"] == a2["A"] and a1["B"] == a2["B"] and a1["root"] == a2["root"]
    print(f"cell batch 2^14 openings: gpu {gpu_s * 1e3:.1f} ms (host buffers, device ms {a1['stage_ms']['total']:.2f}), cpu oracle {cpu_s * 1e3:.0f} ms")
    bad = bytearray(cells); bad[2048 * 7777 + 32 * 13 + 30] ^= 4
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == (0, False)
    ctx.close(); octx.close()


def test_gpu_random_differential(gpu_ctx, oracle_ctx):
    ps.check_random_differential(gpu_ctx, oracle_ctx)


def test_gpu_full_size_2p20_properties(gpu_lib, oracle_lib):
    """BASELINE.json config[3] size (n = 2^20) through size-independent properties: a valid batch is accepted and
    its pairing inputs satisfy A + tau*B = O (pairing-free check with the known test tau); one swapped
    pair of proofs is rejected and breaks that relation; two virtual shards give byte-identical A, B, root."""
    n, seed = 1 << 20, 0x4B5A4703
    ctx = gpu_lib.test_context(devices=[0, 0], n_max=n)
    C, Z, Y, PI = ctx.synth_instance(seed, 0, n)
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)            # two slots on device 0
    a2 = ctx.last_artifacts()
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(a2["A"], a2["B"]) == 1
    one = gpu_lib.test_context(n_max=n)
    assert one.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)            # one slot
    a1 = one.last_artifacts()
    assert a1["A"] == a2["A"] and a1["B"] == a2["B"] and a1["root"] == a2["root"] and a1["sum_ry"] == a2["sum_ry"]
    i, j = 123456, 987654
    PIb = bytearray(PI)
    PIb[48 * i:48 * i + 48], PIb[48 * j:48 * j + 48] = PI[48 * j:48 * j + 48], PI[48 * i:48 * i + 48]
    assert one.verify_kzg_proof_batch(C, Z, Y, bytes(PIb), n) == (0, False)
    ab = one.last_artifacts()
    assert oracle_lib.lib.kzgb_oracle_tau_shortcut(ab["A"], ab["B"]) == 0
    # a slice of the device generator's stream equals the oracle generator's bytes
    octx = oracle_lib.test_context()
    # byte-exact diff of every artefact against the oracle at the benchmarked size (c = 16): ~25 s of CPU on 16 threads
    assert one.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    a1 = one.last_artifacts()
    assert octx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    ao = octx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert a1[key] == ao[key], key
    assert a2["A"] == ao["A"] and a2["B"] == ao["B"] and a2["sum_ry"] == ao["sum_ry"]      # two shards vs oracle
    assert octx.synth_instance(seed, 777777, 32) == tuple(x[w * 777777:w * (777777 + 32)] for x, w in ((C, 48), (Z, 32), (Y, 32), (PI, 48)))
    # MSM linearity at full size: MSM(k) + MSM(k') == MSM(k + k') over 2^20 points
    rc, aff, st = one.g1_decompress_batch(C)
    assert rc == 0 and not any(st)
    import numpy as np
    rng = np.random.default_rng(3)
    k1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); k1[:, 0] &= 0x1F
    k2 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); k2[:, 0] &= 0x1F
    s_ = k1.astype(np.uint16)[:, ::-1] + k2.astype(np.uint16)[:, ::-1]          # little-endian byte order for the carry walk
    carry = np.zeros(n, dtype=np.uint16)
    out = np.zeros((n, 32), dtype=np.uint8)
    for b_ in range(32):
        v = s_[:, b_] + carry
        out[:, b_] = (v & 0xFF).astype(np.uint8)
        carry = v >> 8
    ksum = out[:, ::-1].copy()
    r1 = one.g1_msm(aff, k1.tobytes(), 255)
    r2 = one.g1_msm(aff, k2.tobytes(), 255)
    r3 = one.g1_msm(aff, ksum.tobytes(), 255)
    assert r1[0] == r2[0] == r3[0] == 0
    unit = (1).to_bytes(32, "big")
    assert one.g1_msm(r1[1] + r2[1], unit + unit, 255) == (0, r3[1]) and r3[1] != bytes(96)
    one.close(); ctx.close(); octx.close()


def test_gpu_in_process_multi_device_equals_single(gpu_lib, oracle_lib):
    """The in-process multi-device path of the C ABI (kzgb_ctx_create(devices = {0..G-1}) -> verify_kzg_proof_batch,
    BASELINE.json:5 "combined on the host, no NCCL") on every physical device of the box: verdict, A, B, sum_ry and
    root equal the one-device run and the oracle; a planted wrong proof in the LAST shard is rejected; an
    off-subgroup point in a middle shard is BADARGS with the oracle's count."""
    import torch
    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip("needs at least 2 physical GPUs (run with gpurun --gpus 2)")
    n, seed = (1 << 17) + 4321, 0x4B5A4731
    one = gpu_lib.test_context(devices=[0], n_max=n)
    octx = oracle_lib.test_context()
    C, Z, Y, PI = one.synth_instance(seed, 0, n)
    assert one.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    a1 = one.last_artifacts()
    assert octx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    ao = octx.last_artifacts()
    bad = bytearray(PI)
    bad[48 * (n - 2):48 * (n - 1)], bad[48 * (n - 1):48 * n] = PI[48 * (n - 1):48 * n], PI[48 * (n - 2):48 * (n - 1)]
    off = bytearray(C)
    off[48 * (n // 2):48 * (n // 2 + 1)] = b.g1_compress((0, 2))
    counts = sorted({2, G} | ({4} if G >= 4 else set()))
    for g in counts:
        ctx = gpu_lib.test_context(devices=list(range(g)), n_max=n)
        # the library leaves the caller's current device alone (PyTorch allocates on cudaGetDevice())
        assert torch.empty(1, device="cuda").device.index == torch.cuda.current_device() == 0
        for _ in range(2):                                      # second call: every slot's workspaces are reused
            assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
            ag = ctx.last_artifacts()
            for key in ("A", "B", "sum_ry", "root"):
                assert ag[key] == a1[key] == ao[key], (g, key)
        assert ctx.verify_kzg_proof_batch(C, Z, Y, bytes(bad), n) == (0, False)
        assert ctx.verify_kzg_proof_batch(bytes(off), Z, Y, PI, n) == (1, False)
        assert ctx.last_artifacts()["n_bad_points"] == 1
        assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
        ctx.close()
    one.close(); octx.close()


def test_gpu_cell_batch_multi_device(gpu_lib, oracle_lib):
    """BASELINE.json config[4] "on 8xB200": the 2^14 openings sharded over every physical device of the box (in-process
    context over devices 0..G-1): verdict, pairing inputs A, B and the root equal the oracle's and the one-device run; a
    tampered evaluation in the LAST shard is rejected; a proof outside G1 in a middle shard is BADARGS with count 1."""
    import torch
    from tests.test_cells_oracle import synth_cells
    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip("needs at least 2 physical GPUs (run with gpurun --gpus 2)")
    g1, g2 = oracle_lib.synth_setup(64, 65)
    octx = oracle_lib.context(g1, g2)
    comms, ci, xi, cells, proofs = synth_cells(oracle_lib, 0x4B5A4704, 128, 128, 4096)
    m = len(ci)
    assert octx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    ao = octx.last_artifacts()
    bad = bytearray(cells); bad[2048 * (m - 3) + 32 * 13 + 30] ^= 4
    off = bytearray(proofs); off[48 * (m // 2):48 * (m // 2 + 1)] = b.g1_compress((0, 2))
    for g in sorted({1, 2, G}):
        ctx = gpu_lib.context(g1, g2, devices=list(range(g)), n_max=1 << 15)
        for _ in range(2):
            assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
            a = ctx.last_artifacts()
            assert a["A"] == ao["A"] and a["B"] == ao["B"] and a["root"] == ao["root"], g
        assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == (0, False)
        assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, bytes(off)) == (1, False)
        assert ctx.last_artifacts()["n_bad_points"] == 1
        # a small batch on the same context uses fewer devices
        small = synth_cells(oracle_lib, 0x4B5A4727, 3, 40, 256)
        assert ctx.verify_cell_kzg_proof_batch(*small) == octx.verify_cell_kzg_proof_batch(*small) == (0, True)
        ctx.close()
    octx.close()
