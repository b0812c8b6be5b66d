"""Cell batch (BASELINE.json config[4]) on the CPU oracle vs the pure-Python model."""
import ctypes

import pytest

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k


def test_committed_cell_setup_equals_oracle_generator(oracle_lib):
    from kzg_batch_verification_scheme_b200.api import test_setup
    assert test_setup(cells=True) == oracle_lib.synth_setup(64, 65)


def synth_cells(oracle_lib, seed, n_blobs, cells_per_blob, ncoef, threads=0):
    m = n_blobs * cells_per_blob
    comms = ctypes.create_string_buffer(48 * n_blobs)
    ci, xi = (ctypes.c_uint32 * m)(), (ctypes.c_uint32 * m)()
    cells, proofs = ctypes.create_string_buffer(2048 * m), ctypes.create_string_buffer(48 * m)
    f = oracle_lib.lib.kzgb_oracle_synth_cells
    f.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert f(seed, n_blobs, cells_per_blob, ncoef, comms, ci, xi, cells, proofs, threads) == 0
    return comms.raw, list(ci), list(xi), cells.raw, proofs.raw


@pytest.fixture(scope="module")
def cell_ctx(oracle_lib):
    g1, g2 = oracle_lib.synth_setup(64, 65)
    ctx = oracle_lib.context(g1, g2)
    yield ctx
    ctx.close()


def test_cell_generator_and_artifacts_match_model(oracle_lib, cell_ctx):
    seed = 0x4B5A4704
    got = synth_cells(oracle_lib, seed, 2, 3, 70)
    want = k.gen_cells(seed, 2, 3, ncoef=70)
    assert got == want
    comms, ci, xi, cells, proofs = got
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    art = cell_ctx.last_artifacts()
    m = k.cell_batch_artifacts(comms, ci, xi, cells, proofs)
    assert m["ret"] == 0 and k.cell_verdict_tau_shortcut(m)
    assert art["A"] == b.g1_affine_bytes(m["A"]) and art["B"] == b.g1_affine_bytes(m["B"]) and art["root"] == m["root"]


def test_cell_batch_verdicts(oracle_lib, cell_ctx):
    comms, ci, xi, cells, proofs = synth_cells(oracle_lib, 0x4B5A4714, 4, 8, 4096)
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (0, True)
    # any subset / order of openings verifies
    sel = [5, 30, 2, 17, 17]
    sub = lambda blob, w: b"".join(blob[w * i:w * i + w] for i in sel)    # noqa: E731
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, [ci[i] for i in sel], [xi[i] for i in sel], sub(cells, 2048), sub(proofs, 48)) == (0, True)
    # one evaluation tampered
    bad = bytearray(cells); bad[2048 * 9 + 32 * 5 + 31] ^= 1
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == (0, False)
    # wrong cell index for a valid cell
    xi2 = list(xi); xi2[3] = (xi2[3] + 1) % 128
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi2, cells, proofs) == (0, False)
    # wrong commitment index
    ci2 = list(ci); ci2[0] = 1
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci2, xi, cells, proofs) == (0, False)
    # proofs swapped
    pr2 = proofs[48:96] + proofs[:48] + proofs[96:]
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, pr2) == (0, False)
    # malformed: evaluation >= r, cell index out of range, commitment index out of range, off-subgroup proof
    bad = bytearray(cells); bad[0:32] = b"\xff" * 32
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi, bytes(bad), proofs) == (1, False)
    xi3 = list(xi); xi3[1] = 128
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi3, cells, proofs) == (1, False)
    ci3 = list(ci); ci3[1] = 4
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci3, xi, cells, proofs) == (1, False)
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, b.g1_compress((0, 2)) + proofs[48:]) == (1, False)
    assert cell_ctx.verify_cell_kzg_proof_batch(comms, [], [], b"", b"") == (1, False)


def test_cell_batch_needs_extended_setup(oracle_lib):
    ctx = oracle_lib.test_context()           # 1 G1 + 2 G2 points only
    comms, ci, xi, cells, proofs = synth_cells(oracle_lib, 1, 1, 1, 64)
    assert ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs) == (1, False)
    ctx.close()
