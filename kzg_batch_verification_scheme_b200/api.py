"""ctypes binding of the kzgb200.h C ABI (include/kzgb200.h) -- the host-side mirror of the boundary.

The upstream reference ships no code (only /root/reference/LICENSE:1-201), so the interface mirrored
here is the one BASELINE.json:5 names: ``verify_kzg_proof`` / ``verify_kzg_proof_batch`` behind a
thin C ABI.  `KzgLib` binds *any* shared library exporting that ABI; `load()` returns the CUDA
product library and raises if it has not been built -- there is no CPU fallback in this package.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

KZGB_OK, KZGB_BADARGS, KZGB_ERROR, KZGB_MALLOC = 0, 1, 2, 3
ST_OK, ST_BAD_FLAGS, ST_X_GE_P, ST_NOT_ON_CURVE, ST_NOT_IN_G1 = 0, 1, 2, 3, 4
CHUNK = 128
PARTIAL_BYTES = 320
N_TERMS = 66
TERMS_BYTES = N_TERMS * 192 + 32
N_STAGES = 10
STAGE_NAMES = ["h2d", "decompress", "hash", "root_host", "challenges", "msm_sort", "msm_accumulate",
               "msm_reduce", "pairing", "total"]

OP = dict(FP_MUL=1, FP_SQR=2, FP_ADD=3, FP_SUB=4, FP_INV=5, FP_SQRT_CAND=6, FR_MUL=7, FR_ADD=8, G1_ADD=9,
          G1_DBL=10, G1_MUL=11, G1_MUL_XSQ=12, FP12_MUL=13, FP12_FROB1=14, FP12_FROB2=15, FP12_INV=16,
          FINAL_EXP=17, MILLER_FE=18, SHA256_64=19,
          FPD_MUL=20, FPD_SQR_CHAIN=21)      # 20, 21: FP64-limb multiplier of tools/microbench/fpd.cuh -- tests/emu only, not in the product
OP_SIZES = {1: (96, 48), 2: (48, 48), 3: (96, 48), 4: (96, 48), 5: (48, 48), 6: (48, 48), 7: (64, 32), 8: (64, 32),
            9: (192, 96), 10: (96, 96), 11: (128, 96), 12: (96, 96), 13: (1152, 576), 14: (576, 576),
            15: (576, 576), 16: (576, 576), 17: (576, 576), 18: (192, 576), 19: (64, 32), 20: (96, 48), 21: (48, 48)}


class Artifacts(C.Structure):
    _fields_ = [("S1", C.c_uint8 * 96), ("S2", C.c_uint8 * 96), ("S3", C.c_uint8 * 96), ("A", C.c_uint8 * 96),
                ("B", C.c_uint8 * 96), ("sum_ry", C.c_uint8 * 32), ("root", C.c_uint8 * 32), ("n", C.c_uint64),
                ("n_bad_points", C.c_uint32), ("n_bad_scalars", C.c_uint32), ("stage_ms", C.c_float * N_STAGES)]

    def as_dict(self):
        d = {k: bytes(getattr(self, k)) for k in ("S1", "S2", "S3", "A", "B", "sum_ry", "root")}
        d.update(n=self.n, n_bad_points=self.n_bad_points, n_bad_scalars=self.n_bad_scalars,
                 stage_ms={STAGE_NAMES[i]: self.stage_ms[i] for i in range(N_STAGES)})
        return d


class KzgError(RuntimeError):
    pass


_u8p = C.c_void_p      # raw addresses: host bytes objects or device pointers (ints)


def _ptr(b):
    """bytes / bytearray / ctypes array / int address -> c_void_p."""
    if b is None:
        return None
    if isinstance(b, int):
        return C.c_void_p(b)
    if isinstance(b, bytes):
        return C.cast(C.c_char_p(b), C.c_void_p)
    if isinstance(b, (bytearray, memoryview)):
        return C.cast((C.c_uint8 * len(b)).from_buffer(b), C.c_void_p)
    return C.cast(b, C.c_void_p)


class KzgLib:
    """Binds one shared library that exports include/kzgb200.h."""

    def __init__(self, path):
        self.path = str(path)
        self.lib = lib = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
        sig = {
            "kzgb_ctx_create": [C.POINTER(vp), vp, sz, vp, sz, vp, i32, sz],
            "verify_kzg_proof": [C.POINTER(C.c_bool), vp, vp, vp, vp, vp],
            "verify_kzg_proof_batch": [C.POINTER(C.c_bool), vp, vp, vp, vp, sz, vp],
            "verify_kzg_proof_batch_device": [C.POINTER(C.c_bool), vp, vp, vp, vp, sz, vp, vp],
            "verify_cell_kzg_proof_batch": [C.POINTER(C.c_bool), vp, sz, vp, vp, vp, vp, sz, vp],
            "kzgb_pipeline_init": [vp, i32],
            "verify_kzg_proof_batch_submit": [C.POINTER(u64), vp, vp, vp, vp, sz, i32, vp],
            "verify_kzg_proof_batch_wait": [C.POINTER(C.c_bool), u64, vp],
            "kzgb_shard_phase1": [vp, i32, vp, vp, vp, vp, sz, i32, vp, vp, C.POINTER(C.c_uint32)],
            "kzgb_fs_root": [vp, vp, sz, u64],
            "kzgb_shard_phase2": [vp, i32, vp, u64, vp, vp],
            "kzgb_combine_verify": [vp, vp, i32, C.POINTER(C.c_bool)],
            "kzgb_shard_phase2_terms": [vp, i32, vp, u64, vp, vp],
            "kzgb_shard_finish": [vp, i32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
            "kzgb_combine_verify_terms": [vp, vp, i32, C.POINTER(C.c_bool)],
            "kzgb_g1_decompress_batch": [vp, vp, vp, sz, vp],
            "kzgb_fs_challenges": [vp, vp, vp, vp, vp, vp, sz, vp],
            "kzgb_g1_msm": [vp, vp, vp, sz, i32, vp],
            "kzgb_g1_msm_times": [C.POINTER(C.c_float * 4), vp],
            "kzgb_pairing_check": [C.POINTER(C.c_bool), vp, vp, vp],
            "kzgb_last_artifacts": [vp, C.POINTER(Artifacts)],
            "kzgb_synth_instance": [vp, u64, u64, sz, vp, vp, vp, vp, i32],
            "kzgb_debug_op": [vp, i32, vp, vp, sz],
            "kzgb_imad_peak": [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)],
            "kzgb_imad32_peak": [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)],
            "kzgb_last_stage_ms": [vp, C.POINTER(C.c_float * N_STAGES)],
            "kzgb_set_subgroup_batch_min": [vp, sz],
            "verify_blob_kzg_proof_batch": [C.POINTER(C.c_bool), vp, vp, vp, sz, vp],
            "kzgb_blob_challenges_evals": [vp, vp, vp, vp, sz, vp],
            "kzgb_blob_eval": [vp, vp, vp, sz, vp],
            "verify_kzg_proof_batch_eip4844": [C.POINTER(C.c_bool), vp, vp, vp, vp, sz, vp],
            "verify_blob_kzg_proof_batch_eip4844": [C.POINTER(C.c_bool), vp, vp, vp, sz, vp],
            "kzgb_blob_challenges_evals_eip4844": [vp, vp, vp, vp, sz, vp],
            "kzgb_load_trusted_setup_file": [C.POINTER(vp), C.c_char_p, vp, i32, sz],
        }
        for name, args in sig.items():
            f = getattr(lib, name)
            f.argtypes, f.restype = args, i32
        lib.kzgb_ctx_free.argtypes, lib.kzgb_ctx_free.restype = [vp], None
        lib.kzgb_launch_count.argtypes, lib.kzgb_launch_count.restype = [vp], u64
        lib.kzgb_set_threads.argtypes, lib.kzgb_set_threads.restype = [vp, i32], i32
        lib.kzgb_version.argtypes, lib.kzgb_version.restype = [], C.c_char_p
        if hasattr(lib, "kzgb_synth_setup"):                 # oracle library only (include/kzgb200_testing.h)
            lib.kzgb_synth_setup.argtypes, lib.kzgb_synth_setup.restype = [vp, sz, vp, sz], i32

    EXPORTS = ["kzgb_ctx_create", "kzgb_ctx_free", "verify_kzg_proof", "verify_kzg_proof_batch",
               "verify_kzg_proof_batch_device", "verify_cell_kzg_proof_batch", "kzgb_pipeline_init", "verify_kzg_proof_batch_submit",
               "verify_kzg_proof_batch_wait", "kzgb_shard_phase1", "kzgb_fs_root", "kzgb_shard_phase2",
               "kzgb_combine_verify", "kzgb_shard_phase2_terms", "kzgb_shard_finish", "kzgb_combine_verify_terms", "kzgb_g1_decompress_batch", "kzgb_fs_challenges", "kzgb_g1_msm",
               "kzgb_g1_msm_times", "kzgb_pairing_check", "kzgb_last_artifacts", "kzgb_synth_instance",
               "kzgb_debug_op", "kzgb_imad_peak", "kzgb_imad32_peak", "kzgb_last_stage_ms", "kzgb_launch_count", "kzgb_set_threads",
               "kzgb_set_subgroup_batch_min", "verify_blob_kzg_proof_batch", "kzgb_blob_challenges_evals", "kzgb_blob_eval", "kzgb_version",
               "verify_kzg_proof_batch_eip4844", "verify_blob_kzg_proof_batch_eip4844", "kzgb_blob_challenges_evals_eip4844",
               "kzgb_load_trusted_setup_file"]

    def version(self) -> str:
        return self.lib.kzgb_version().decode()

    def synth_setup(self, n1=1, n2=2):
        """INSECURE test setup (known tau) -- exported by the oracle library only (include/kzgb200_testing.h)."""
        if not hasattr(self.lib, "kzgb_synth_setup"):
            raise KzgError("kzgb_synth_setup is test infrastructure of the oracle library; the product does not export it")
        g1, g2 = C.create_string_buffer(48 * n1), C.create_string_buffer(96 * n2)
        rc = self.lib.kzgb_synth_setup(_ptr(g1), n1, _ptr(g2), n2)
        if rc:
            raise KzgError(f"kzgb_synth_setup -> {rc}")
        return g1.raw, g2.raw

    def context(self, g1_monomial, g2_monomial, devices=None, n_max=1 << 16):
        """Verifier over the caller's trusted setup (compressed [tau^i]G1, [tau^i]G2).  The setup is mandatory:
        a verifier is only as trustworthy as its setup, so there is no default."""
        if g1_monomial is None or g2_monomial is None:
            raise KzgError("a trusted setup is required (g1_monomial, g2_monomial); for tests and benchmarks use "
                           "test_context(), whose tau is public")
        return Context(self, g1_monomial, g2_monomial, devices, n_max)

    def context_from_file(self, path, devices=None, n_max=1 << 16):
        """Verifier over a c-kzg-4844 `trusted_setup.txt` (kzgb_load_trusted_setup_file)."""
        ctx = Context.__new__(Context)
        ctx.klib, ctx.lib = self, self.lib
        nd = len(devices) if devices is not None else 0
        devs = (C.c_int * nd)(*devices) if nd else None
        ctx.n_devices = max(nd, 1)
        h = C.c_void_p()
        rc = self.lib.kzgb_load_trusted_setup_file(C.byref(h), str(path).encode(), C.cast(devs, C.c_void_p) if devs else None, nd, n_max)
        if rc:
            raise KzgError(f"kzgb_load_trusted_setup_file -> {rc}")
        ctx.h = h
        return ctx

    def test_context(self, devices=None, n_max=1 << 16, cells=False):
        """INSECURE: context over the repository's test setup, whose tau is derivable by anyone
        (SHA256("kzgb200/insecure-test-tau") mod r) -- forged proofs verify against it.  Tests and bench only.
        cells=True: the extended setup ([tau^j]G1 j < 64, [tau^j]G2 j <= 64) the cell batch needs."""
        g1, g2 = test_setup(cells)
        return Context(self, g1, g2, devices, n_max)


class Context:
    """Owns one kzgb_ctx.  Methods mirror the C entry points one-to-one (same names, same argument meaning)."""

    def __init__(self, klib: KzgLib, g1_monomial, g2_monomial, devices, n_max):
        self.klib, self.lib = klib, klib.lib
        if g1_monomial is None or g2_monomial is None:
            raise KzgError("a trusted setup is required (see KzgLib.test_context for the insecure test setup)")
        devs = None
        nd = 0
        if devices is not None:
            nd = len(devices)
            devs = (C.c_int * nd)(*devices)
        self.n_devices = max(nd, 1)
        h = C.c_void_p()
        rc = self.lib.kzgb_ctx_create(C.byref(h), _ptr(g1_monomial), len(g1_monomial) // 48, _ptr(g2_monomial),
                                      len(g2_monomial) // 96, C.cast(devs, C.c_void_p) if devs else None, nd, n_max)
        if rc:
            raise KzgError(f"kzgb_ctx_create -> {rc}")
        self.h = h

    def close(self):
        if self.h:
            self.lib.kzgb_ctx_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the two entry points BJ:5 names
    def verify_kzg_proof(self, Cb, z, y, pi):
        ok = C.c_bool(False)
        rc = self.lib.verify_kzg_proof(C.byref(ok), _ptr(Cb), _ptr(z), _ptr(y), _ptr(pi), self.h)
        return rc, bool(ok.value)

    def verify_kzg_proof_batch(self, Cb, z, y, pi, n):
        ok = C.c_bool(False)
        rc = self.lib.verify_kzg_proof_batch(C.byref(ok), _ptr(Cb), _ptr(z), _ptr(y), _ptr(pi), n, self.h)
        return rc, bool(ok.value)

    def verify_kzg_proof_batch_device(self, dC, dz, dy, dpi, n, stream=0):
        ok = C.c_bool(False)
        rc = self.lib.verify_kzg_proof_batch_device(C.byref(ok), _ptr(dC), _ptr(dz), _ptr(dy), _ptr(dpi), n, self.h,
                                                    C.c_void_p(stream))
        return rc, bool(ok.value)

    def pipeline_init(self, depth: int) -> int:
        """Allocate `depth` workspaces for verify_kzg_proof_batch_submit / _wait (batches in flight on device 0)."""
        return int(self.lib.kzgb_pipeline_init(self.h, depth))

    def verify_kzg_proof_batch_submit(self, Cb, z, y, pi, n, on_device=False):
        """(rc, ticket).  The buffers must stay alive and untouched until verify_kzg_proof_batch_wait(ticket) returns."""
        t = C.c_uint64(0)
        rc = self.lib.verify_kzg_proof_batch_submit(C.byref(t), _ptr(Cb), _ptr(z), _ptr(y), _ptr(pi), n, int(on_device), self.h)
        return rc, int(t.value)

    def verify_kzg_proof_batch_wait(self, ticket: int):
        ok = C.c_bool(False)
        rc = self.lib.verify_kzg_proof_batch_wait(C.byref(ok), ticket, self.h)
        return rc, bool(ok.value)

    def verify_cell_kzg_proof_batch(self, commitments: bytes, commitment_indices, cell_indices, cells: bytes, proofs: bytes):
        """Cell batch (PeerDAS-shaped): index lists are sequences of ints; returns (rc, ok)."""
        m = len(commitment_indices)
        ci = (C.c_uint32 * m)(*commitment_indices)
        xi = (C.c_uint32 * m)(*cell_indices)
        ok = C.c_bool(False)
        rc = self.lib.verify_cell_kzg_proof_batch(C.byref(ok), _ptr(commitments), len(commitments) // 48, C.cast(ci, C.c_void_p),
                                                  C.cast(xi, C.c_void_p), _ptr(cells), _ptr(proofs), m, self.h)
        return rc, bool(ok.value)

    # ---- shard level
    def shard_phase1(self, slot, Cb, z, y, pi, n_local, on_device=False, stream=0):
        dig = C.create_string_buffer(32 * ((n_local + CHUNK - 1) // CHUNK))
        nbad = C.c_uint32(0)
        rc = self.lib.kzgb_shard_phase1(self.h, slot, _ptr(Cb), _ptr(z), _ptr(y), _ptr(pi), n_local, int(on_device),
                                        C.c_void_p(stream), _ptr(dig), C.byref(nbad))
        return rc, dig.raw, nbad.value

    def fs_root(self, digests: bytes, n_total: int) -> bytes:
        root = C.create_string_buffer(32)
        rc = self.lib.kzgb_fs_root(_ptr(root), _ptr(digests), len(digests) // 32, n_total)
        if rc:
            raise KzgError(f"kzgb_fs_root -> {rc}")
        return root.raw

    def shard_phase2(self, slot, root: bytes, global_offset: int, stream=0):
        out = C.create_string_buffer(PARTIAL_BYTES)
        rc = self.lib.kzgb_shard_phase2(self.h, slot, _ptr(root), global_offset, C.c_void_p(stream), _ptr(out))
        return rc, out.raw

    def combine_verify(self, partials: bytes):
        ok = C.c_bool(False)
        rc = self.lib.kzgb_combine_verify(self.h, _ptr(partials), len(partials) // PARTIAL_BYTES, C.byref(ok))
        return rc, bool(ok.value)

    def shard_phase2_terms(self, slot, root: bytes, global_offset: int, stream=0, out=None):
        """(rc, record): the shard's 66 pairing terms + sum r_i y_i (TERMS_BYTES); `out` = writable buffer to fill in place."""
        buf = out if out is not None else C.create_string_buffer(TERMS_BYTES)
        rc = self.lib.kzgb_shard_phase2_terms(self.h, slot, _ptr(root), global_offset, C.c_void_p(stream), _ptr(buf))
        return rc, (buf if out is not None else buf.raw)

    def shard_finish(self, slot=0):
        """(rc, n_bad_points, n_bad_scalars) of the shard's input validation; rc = 1 if any element is malformed."""
        bp, bs = C.c_uint32(0), C.c_uint32(0)
        rc = self.lib.kzgb_shard_finish(self.h, slot, C.byref(bp), C.byref(bs))
        return rc, bp.value, bs.value

    def combine_verify_terms(self, terms, n_shards=None):
        n_shards = len(terms) // TERMS_BYTES if n_shards is None else n_shards
        ok = C.c_bool(False)
        rc = self.lib.kzgb_combine_verify_terms(self.h, _ptr(terms), n_shards, C.byref(ok))
        return rc, bool(ok.value)

    # ---- stage exports
    def g1_decompress_batch(self, data: bytes):
        m = len(data) // 48
        aff, st = C.create_string_buffer(96 * m), C.create_string_buffer(m)
        rc = self.lib.kzgb_g1_decompress_batch(_ptr(aff), _ptr(st), _ptr(data), m, self.h)
        return rc, aff.raw, st.raw

    def fs_challenges(self, Cb, z, y, pi, n):
        root, r = C.create_string_buffer(32), C.create_string_buffer(16 * n)
        rc = self.lib.kzgb_fs_challenges(_ptr(root), _ptr(r), _ptr(Cb), _ptr(z), _ptr(y), _ptr(pi), n, self.h)
        return rc, root.raw, r.raw

    def g1_msm(self, points_affine: bytes, scalars: bytes, nbits=255):
        m = len(points_affine) // 96
        out = C.create_string_buffer(96)
        rc = self.lib.kzgb_g1_msm(_ptr(out), _ptr(points_affine), _ptr(scalars), m, nbits, self.h)
        return rc, out.raw

    def g1_msm_times(self):
        ms = (C.c_float * 4)()
        self.lib.kzgb_g1_msm_times(C.byref(ms), self.h)
        return list(ms)

    def pairing_check(self, A: bytes, B: bytes):
        ok = C.c_bool(False)
        rc = self.lib.kzgb_pairing_check(C.byref(ok), _ptr(A), _ptr(B), self.h)
        return rc, bool(ok.value)

    def last_artifacts(self) -> dict:
        a = Artifacts()
        rc = self.lib.kzgb_last_artifacts(self.h, C.byref(a))
        if rc:
            raise KzgError(f"kzgb_last_artifacts -> {rc}")
        return a.as_dict()

    def synth_instance(self, seed, offset, n, device_ptrs=None):
        """Host bytes (C, z, y, pi) or, with device_ptrs=(dC,dz,dy,dpi), fills those device buffers."""
        if device_ptrs is not None:
            dC, dz, dy, dpi = device_ptrs
            rc = self.lib.kzgb_synth_instance(self.h, seed, offset, n, _ptr(dC), _ptr(dz), _ptr(dy), _ptr(dpi), 1)
            if rc:
                raise KzgError(f"kzgb_synth_instance -> {rc}")
            return None
        bufs = [C.create_string_buffer(s * n) for s in (48, 32, 32, 48)]
        rc = self.lib.kzgb_synth_instance(self.h, seed, offset, n, *[_ptr(b) for b in bufs], 0)
        if rc:
            raise KzgError(f"kzgb_synth_instance -> {rc}")
        return tuple(b.raw for b in bufs)

    def debug_op(self, op, data: bytes):
        op = OP[op] if isinstance(op, str) else op
        isz, osz = OP_SIZES[op]
        cnt = len(data) // isz
        out = C.create_string_buffer(osz * cnt)
        rc = self.lib.kzgb_debug_op(self.h, op, _ptr(data), _ptr(out), cnt)
        return rc, out.raw

    def imad_peak(self, wide=True):
        """(ops/s, ms) of the multiply-add microbenchmark: wide=True -> IMAD.WIDE.U32.X carry chains."""
        v, ms = C.c_double(0), C.c_double(0)
        fn = self.lib.kzgb_imad_peak if wide else self.lib.kzgb_imad32_peak
        rc = fn(self.h, C.byref(v), C.byref(ms))
        if rc:
            raise KzgError(f"kzgb_imad_peak -> {rc}")
        return v.value, ms.value

    def last_stage_ms(self) -> dict:
        ms = (C.c_float * N_STAGES)()
        rc = self.lib.kzgb_last_stage_ms(self.h, C.byref(ms))
        if rc:
            raise KzgError(f"kzgb_last_stage_ms -> {rc}")
        return {STAGE_NAMES[i]: ms[i] for i in range(N_STAGES)}

    def launch_count(self) -> int:
        return int(self.lib.kzgb_launch_count(self.h))

    def set_threads(self, n: int) -> int:
        return int(self.lib.kzgb_set_threads(self.h, n))

    def verify_blob_kzg_proof_batch(self, blobs, commitments: bytes, proofs: bytes, m=None):
        """(rc, ok): m blobs of 4096 x 32 B (bytes or a raw host address), m commitments, m proofs
        (include/kzgb200.h "Blob batch")."""
        m = len(commitments) // 48 if m is None else m
        assert isinstance(blobs, int) or len(blobs) == 131072 * m
        assert len(proofs) == 48 * m
        ok = C.c_bool(False)
        rc = self.lib.verify_blob_kzg_proof_batch(C.byref(ok), _ptr(blobs), _ptr(commitments), _ptr(proofs), m, self.h)
        return rc, bool(ok.value)

    def blob_challenges_evals(self, blobs: bytes, commitments: bytes):
        m = len(commitments) // 48
        z, y = C.create_string_buffer(32 * m), C.create_string_buffer(32 * m)
        rc = self.lib.kzgb_blob_challenges_evals(z, y, _ptr(blobs), _ptr(commitments), m, self.h)
        return rc, z.raw, y.raw

    def blob_eval(self, blobs: bytes, z: bytes):
        m = len(z) // 32
        y = C.create_string_buffer(32 * m)
        rc = self.lib.kzgb_blob_eval(y, _ptr(blobs), _ptr(z), m, self.h)
        return rc, y.raw

    # ---- EIP-4844 / c-kzg-4844 transcript mode
    def verify_kzg_proof_batch_eip4844(self, Cb, z, y, pi, n):
        ok = C.c_bool(False)
        rc = self.lib.verify_kzg_proof_batch_eip4844(C.byref(ok), _ptr(Cb), _ptr(z), _ptr(y), _ptr(pi), n, self.h)
        return rc, bool(ok.value)

    def verify_blob_kzg_proof_batch_eip4844(self, blobs, commitments: bytes, proofs: bytes, m=None):
        m = len(commitments) // 48 if m is None else m
        ok = C.c_bool(False)
        rc = self.lib.verify_blob_kzg_proof_batch_eip4844(C.byref(ok), _ptr(blobs), _ptr(commitments), _ptr(proofs), m, self.h)
        return rc, bool(ok.value)

    def blob_challenges_evals_eip4844(self, blobs: bytes, commitments: bytes):
        m = len(commitments) // 48
        z, y = C.create_string_buffer(32 * m), C.create_string_buffer(32 * m)
        rc = self.lib.kzgb_blob_challenges_evals_eip4844(z, y, _ptr(blobs), _ptr(commitments), m, self.h)
        return rc, z.raw, y.raw

    def set_subgroup_batch_min(self, n_min: int) -> int:
        """Batches of >= n_min proofs use the batched subgroup check (0: always the per-point check)."""
        return int(self.lib.kzgb_set_subgroup_batch_min(self.h, n_min))


PKG_DIR = Path(__file__).resolve().parent


def test_setup(cells=False):
    """The insecure TEST trusted setup (known tau; tools/gen_test_setup.py): ([tau^0]G1, [tau^0..1]G2), or with
    cells=True ([tau^j]G1 j < 64, [tau^j]G2 j <= 64)."""
    if cells:
        blob = (PKG_DIR / "data" / "test_setup_cells.bin").read_bytes()
        return blob[:64 * 48], blob[64 * 48:]
    blob = (PKG_DIR / "data" / "test_setup.bin").read_bytes()
    return blob[:48], blob[48:240]

PRODUCT_LIB = PKG_DIR / "csrc" / "libkzgb200.so"


def load() -> KzgLib:
    """The CUDA product library.  Raises if it is not built: this package has no CPU fallback."""
    if not PRODUCT_LIB.exists():
        raise KzgError(f"{PRODUCT_LIB} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    return KzgLib(PRODUCT_LIB)
