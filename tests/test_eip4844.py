"""EIP-4844 / c-kzg-4844 transcript mode.  CPU: the oracle against the pure-Python restatement of
the consensus-spec functions (oracle/pymodel) -- digest, verdict, pairing inputs, blob challenge and evaluation, the
trusted_setup.txt loader.  GPU: the CUDA library against the oracle.  No c-kzg vector exists offline ("parity unpinned")."""
import random

import pytest

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k


def _affine(pt):
    return b.g1_affine_bytes(pt)


def _check_against_model(ctx, C, Z, Y, PI, n):
    m = k.eip_batch_artifacts(C, Z, Y, PI, n)
    rc, ok = ctx.verify_kzg_proof_batch_eip4844(C, Z, Y, PI, n)
    assert rc == m["ret"]
    if rc:
        return rc, ok
    a = ctx.last_artifacts()
    assert a["root"] == m["root"] and a["A"] == _affine(m["A"]) and a["B"] == _affine(m["B"])
    assert a["sum_ry"] == m["sum_ry"].to_bytes(32, "big")
    assert ok == k.verdict_tau_shortcut(m)
    return rc, ok


def test_oracle_batch_matches_python_restatement(oracle_ctx):
    for n in (1, 2, 5):
        C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A4790 + n, 0, n)
        assert _check_against_model(oracle_ctx, C, Z, Y, PI, n) == (0, True)
        if n >= 2:
            PI2 = PI[48:96] + PI[:48] + PI[96:]
            assert _check_against_model(oracle_ctx, C, Z, Y, PI2, n) == (0, False)
            Z2 = b"\xff" * 32 + Z[32:]
            assert _check_against_model(oracle_ctx, C, Z2, Y, PI, n)[0] == 1
            C2 = b.g1_compress((0, 2)) + C[48:]                    # on the curve, outside G1
            assert _check_against_model(oracle_ctx, C2, Z, Y, PI, n)[0] == 1


def test_oracle_blob_path_matches_python_restatement(oracle_ctx, oracle_lib):
    from tests import parity_suite as ps
    blobs, comms, proofs = ps.synth_blobs(oracle_lib, 0x4B5A4795, 2)
    rc, zs, ys = oracle_ctx.blob_challenges_evals_eip4844(blobs, comms)
    assert rc == 0
    for j in range(2):
        blob = blobs[131072 * j:131072 * (j + 1)]
        z = k.eip_blob_challenge(blob, comms[48 * j:48 * j + 48])
        assert zs[32 * j:32 * j + 32] == z.to_bytes(32, "big")
        assert ys[32 * j:32 * j + 32] == k.eip_blob_eval(blob, z).to_bytes(32, "big")
    # real blobs with their real commitments and proofs made for THIS library's challenge: under the EIP-4844 challenge the
    # proofs open at a different point, so the batch is well-formed and must be rejected; the (z, y) it checks are the
    # EIP-4844 ones
    assert oracle_ctx.verify_blob_kzg_proof_batch_eip4844(blobs, comms, proofs) == (0, False)
    # an evaluation ON the domain returns the blob element itself
    w5 = pow(k.OMEGA_BLOB, k._brp(5, 12), k.R)
    assert k.eip_blob_eval(blobs[:131072], w5) == int.from_bytes(blobs[32 * 5:32 * 6], "big")


def test_blob_proofs_for_the_eip4844_challenge_verify(oracle_ctx, oracle_lib):
    """A blob of known coefficients: commitment and proof built with the test tau for the EIP-4844 challenge z."""
    from tests import parity_suite as ps
    blob, comm, _, coeffs = ps.python_blob(random.Random(21))
    z = k.eip_blob_challenge(blob, comm)
    y = ps.poly_eval(coeffs, z)
    assert k.eip_blob_eval(blob, z) == y
    # proof = [(p(tau) - y) / (tau - z)] G1
    ptau = ps.poly_eval(coeffs, k.TAU)
    proof = b.g1_compress(b.g1_mul((ptau - y) * pow(k.TAU - z, -1, k.R) % k.R, b.G1))
    assert oracle_ctx.verify_blob_kzg_proof_batch_eip4844(blob, comm, proof) == (0, True)
    bad = bytearray(blob); bad[32 * 77 + 31] ^= 1
    assert oracle_ctx.verify_blob_kzg_proof_batch_eip4844(bytes(bad), comm, proof) == (0, False)


def _write_setup_file(path, g1m, g2m, n1, with_monomials):
    lines = [str(n1), str(len(g2m) // 96)]
    lagrange = b.g1_compress(b.G1).hex()                         # placeholder points: the verifier skips this section
    lines += [lagrange] * n1
    lines += [g2m[96 * i:96 * (i + 1)].hex() for i in range(len(g2m) // 96)]
    if with_monomials:
        lines += [g1m[48 * i:48 * (i + 1)].hex() for i in range(n1)]
    path.write_text("\n".join(lines) + "\n")


def test_trusted_setup_file_loader(oracle_lib, tmp_path):
    from kzg_batch_verification_scheme_b200.api import KzgError, test_setup
    g1c, g2c = test_setup(cells=True)
    C, Z, Y, PI = oracle_lib.test_context().synth_instance(0x4B5A4799, 0, 9)
    for with_mono in (False, True):
        p = tmp_path / f"trusted_setup_{int(with_mono)}.txt"
        _write_setup_file(p, g1c, g2c, 64, with_mono)
        ctx = oracle_lib.context_from_file(p)
        assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, 9) == (0, True)
        assert ctx.verify_kzg_proof_batch_eip4844(C, Z, Y, PI, 9) == (0, True)
        ctx.close()
    bad = tmp_path / "bad.txt"
    bad.write_text("64\n65\n" + "zz" * 48 + "\n")
    with pytest.raises(KzgError):
        oracle_lib.context_from_file(bad)
    with pytest.raises(KzgError):
        oracle_lib.context_from_file(tmp_path / "missing.txt")
    assert b.g1_compress(b.G1).hex() == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 7, 128, 1000, 4096])
def test_gpu_eip4844_batch_vs_oracle(gpu_lib, oracle_ctx, n):
    ctx = gpu_lib.test_context(n_max=2 * (4096 + 1))
    C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A47A0 + n, 0, n)
    cases = [(C, Z, Y, PI)]
    if n >= 2:
        cases.append((C, Z, Y, PI[48:96] + PI[:48] + PI[96:]))                       # wrong proof
        cases.append((C, Z[:32 * (n - 1)] + b"\xff" * 32, Y, PI))                     # scalar >= r
        cases.append((C[:48 * (n // 2)] + b.g1_compress((0, 2)) + C[48 * (n // 2 + 1):], Z, Y, PI))   # outside G1
        cases.append((C, Z, Y, PI[:48 * (n - 1)] + b"\xc0" + bytes(47)))              # a proof at infinity: valid encoding
    for (c, z, y, p) in cases:
        got = ctx.verify_kzg_proof_batch_eip4844(c, z, y, p, n)
        want = oracle_ctx.verify_kzg_proof_batch_eip4844(c, z, y, p, n)
        assert got == want, (n, got, want)
        a1, a2 = ctx.last_artifacts(), oracle_ctx.last_artifacts()
        assert a1["n_bad_points"] == a2["n_bad_points"] and a1["n_bad_scalars"] == a2["n_bad_scalars"]
        if got[0] == 0:
            for key in ("A", "B", "sum_ry", "root"):
                assert a1[key] == a2[key], (n, key)
    assert cases[0] and ctx.verify_kzg_proof_batch_eip4844(*cases[0], n) == (0, True)
    # the tree-transcript entry point keeps working on the same context, and a batch too large for this mode is refused
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n) == (0, True)
    big = oracle_ctx.synth_instance(1, 0, 4097)
    assert ctx.verify_kzg_proof_batch_eip4844(*big, 4097) == (1, False)
    ctx.close()


@pytest.mark.gpu
def test_gpu_eip4844_blob_path_vs_oracle(gpu_lib, oracle_ctx, oracle_lib):
    from tests import parity_suite as ps
    ctx = gpu_lib.test_context(n_max=1024)
    blobs, comms, proofs = ps.synth_blobs(oracle_lib, 0x4B5A47B0, 5)
    assert ctx.blob_challenges_evals_eip4844(blobs, comms) == oracle_ctx.blob_challenges_evals_eip4844(blobs, comms)
    assert ctx.verify_blob_kzg_proof_batch_eip4844(blobs, comms, proofs) == oracle_ctx.verify_blob_kzg_proof_batch_eip4844(blobs, comms, proofs) == (0, False)
    blob, comm, _, coeffs = ps.python_blob(random.Random(22))
    z = k.eip_blob_challenge(blob, comm)
    y = ps.poly_eval(coeffs, z)
    proof = b.g1_compress(b.g1_mul((ps.poly_eval(coeffs, k.TAU) - y) * pow(k.TAU - z, -1, k.R) % k.R, b.G1))
    assert ctx.verify_blob_kzg_proof_batch_eip4844(blob, comm, proof) == (0, True)
    assert ctx.blob_challenges_evals_eip4844(blob, comm) == (0, z.to_bytes(32, "big"), y.to_bytes(32, "big"))
    bad = bytearray(blob); bad[:32] = b"\xff" * 32                                     # element >= r
    assert ctx.verify_blob_kzg_proof_batch_eip4844(bytes(bad), comm, proof) == oracle_ctx.verify_blob_kzg_proof_batch_eip4844(bytes(bad), comm, proof) == (1, False)
    ctx.close()


@pytest.mark.gpu
def test_gpu_trusted_setup_file_loader(gpu_lib, oracle_ctx, tmp_path):
    from kzg_batch_verification_scheme_b200.api import test_setup
    g1c, g2c = test_setup(cells=True)
    p = tmp_path / "trusted_setup.txt"
    _write_setup_file(p, g1c, g2c, 64, True)
    ctx = gpu_lib.context_from_file(p, n_max=4096)
    C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A47C0, 0, 300)
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, 300) == (0, True)
    assert ctx.verify_kzg_proof_batch_eip4844(C, Z, Y, PI, 300) == (0, True)
    ctx.close()
