import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ORACLE_LIB = ROOT / "oracle" / "libkzgb_oracle.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure).  Built on demand with its own Makefile."""
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    from kzg_batch_verification_scheme_b200.api import KzgLib
    return KzgLib(ORACLE_LIB)


@pytest.fixture(scope="session")
def oracle_ctx(oracle_lib):
    ctx = oracle_lib.test_context()
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_lib():
    from kzg_batch_verification_scheme_b200.api import load
    return load()


@pytest.fixture(scope="session")
def gpu_ctx(gpu_lib):
    ctx = gpu_lib.test_context(n_max=1 << 16)
    yield ctx
    ctx.close()
