"""Blob batch on the CPU: the oracle against the SPEC restated in Python -- challenge by
hashlib, evaluation by direct polynomial evaluation of a blob built from known coefficients."""
import ctypes as C
import hashlib
import random

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k
from tests import parity_suite as ps
from tests.helpers import R


def test_blob_spec_against_python(oracle_ctx):
    blob, comm, proof, coeffs = ps.python_blob(random.Random(11))
    z = ps.python_blob_challenge(blob, comm)
    y = ps.poly_eval(coeffs, z)
    rc, zs, ys = oracle_ctx.blob_challenges_evals(blob, comm)
    assert rc == 0 and zs == z.to_bytes(32, "big") and ys == y.to_bytes(32, "big")
    assert oracle_ctx.verify_blob_kzg_proof_batch(blob, comm, proof) == (0, True)
    # evaluation at caller-chosen points: a domain point returns the stored evaluation, others p(z)
    pts = [ps.blob_domain_point(0), ps.blob_domain_point(1), ps.blob_domain_point(4095), 0, 1, R - 1, 12345]
    rc, ys = oracle_ctx.blob_eval(blob * len(pts), b"".join(v.to_bytes(32, "big") for v in pts))
    assert rc == 0 and ys == b"".join(ps.poly_eval(coeffs, v).to_bytes(32, "big") for v in pts)
    assert ys[:32] == blob[:32] and ys[32:64] == blob[32:64]


def test_blob_batch_oracle(oracle_lib, oracle_ctx):
    ps.check_blob_batch(oracle_ctx, oracle_ctx, ps.synth_blobs(oracle_lib, 0x4B5A4741, 3))
