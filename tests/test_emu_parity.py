"""CPU-side checks of the CUDA kernels' per-thread logic: the device headers are compiled for the host
(-DKZGB_EMU, tests/emu) and diffed against the oracle with the same suite the GPU tests use."""
import subprocess
from pathlib import Path

import pytest

from tests import parity_suite as ps

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu_ctx(oracle_lib):
    subprocess.run(["make", "-s", "-C", str(ROOT / "tests" / "emu")], check=True)
    from kzg_batch_verification_scheme_b200.api import KzgLib
    lib = KzgLib(ROOT / "tests" / "emu" / "libkzgb_emu.so")
    g1, g2 = oracle_lib.synth_setup(1, 2)
    ctx = lib.context(g1, g2)
    yield ctx
    ctx.close()


def test_emu_field_ops(emu_ctx, oracle_ctx):
    ps.check_field_ops(emu_ctx, oracle_ctx)


def test_emu_fpd_ops(emu_ctx, oracle_ctx):
    ps.check_fpd_ops(emu_ctx, oracle_ctx)


def test_emu_subgroup_batch(emu_ctx, oracle_ctx):
    ps.check_subgroup_batch(emu_ctx, oracle_ctx, n=6)


def test_emu_blob_batch(emu_ctx, oracle_ctx, oracle_lib):
    ps.check_blob_batch(emu_ctx, oracle_ctx, ps.synth_blobs(oracle_lib, 0x4B5A4742, 2))


def test_emu_random_differential_batched_check(emu_ctx, oracle_ctx):
    """The randomised corruptions again with the batched subgroup check switched on in the emulation (the CUDA
    library's default): return codes, verdicts and pairing inputs must still equal the oracle's."""
    assert emu_ctx.set_subgroup_batch_min(2) == 0
    try:
        ps.check_random_differential(emu_ctx, oracle_ctx, sizes=(2, 3, 9), trials=10, pool=12)
    finally:
        emu_ctx.set_subgroup_batch_min(0)


def test_emu_sha_single_block(emu_ctx, oracle_ctx):
    import random
    rnd = random.Random(1)
    data = bytes(rnd.randrange(256) for _ in range(64 * 5))
    assert emu_ctx.debug_op("SHA256_64", data) == oracle_ctx.debug_op("SHA256_64", data)


def test_emu_g1_ops(emu_ctx, oracle_ctx):
    ps.check_g1_ops(emu_ctx, oracle_ctx)


def test_emu_decompress(emu_ctx, oracle_ctx):
    ps.check_decompress(emu_ctx, oracle_ctx, n_valid=6)


def test_emu_tower_and_pairing(emu_ctx, oracle_ctx):
    ps.check_tower_and_pairing(emu_ctx, oracle_ctx)


def test_emu_fs(emu_ctx, oracle_ctx):
    ps.check_fs(emu_ctx, oracle_ctx, sizes=(1, 2, 63, 127, 128, 129, 1025))


def test_emu_msm(emu_ctx, oracle_ctx):
    ps.check_msm(emu_ctx, oracle_ctx, sizes=((1, 255), (2, 255), (7, 255), (64, 255), (64, 128), (300, 128)))


def test_emu_synth(emu_ctx, oracle_ctx):
    ps.check_synth(emu_ctx, oracle_ctx, n=4)


def test_emu_verify(emu_ctx, oracle_ctx, oracle_lib):
    ps.check_verify(emu_ctx, oracle_ctx, oracle_lib, sizes=(1, 2, 9))


def test_emu_degenerate(emu_ctx, oracle_ctx):
    ps.check_degenerate(emu_ctx, oracle_ctx, n=6)


def test_emu_status_classes(emu_ctx, oracle_ctx):
    ps.check_status_classes(emu_ctx, oracle_ctx, n=4)


def test_emu_cell_batch(oracle_lib):
    """Cell batch (config[4]) through the device bodies vs the oracle: verdicts and pairing inputs."""
    from kzg_batch_verification_scheme_b200.api import KzgLib
    from tests.test_cells_oracle import synth_cells
    lib = KzgLib(ROOT / "tests" / "emu" / "libkzgb_emu.so")
    g1, g2 = oracle_lib.synth_setup(64, 65)
    ctx, octx = lib.context(g1, g2), oracle_lib.context(g1, g2)
    ps.check_cell_batch(ctx, octx, synth_cells(oracle_lib, 0x4B5A4724, 2, 3, 200))
    ctx.close(); octx.close()


def test_emu_random_differential(emu_ctx, oracle_ctx):
    ps.check_random_differential(emu_ctx, oracle_ctx, sizes=(1, 2, 3, 9), trials=14, pool=12)
