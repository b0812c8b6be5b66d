"""The C-ABI libraries load and export every symbol include/kzgb200.h declares (no compute without a GPU)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "kzgb200.h").read_text()
    names = set(re.findall(r"\b(kzgb_[a-z0-9_]+|verify_[a-z_]*kzg_proof[a-z_]*)\s*\(", txt))
    return sorted(n for n in names if n not in ("kzgb_ret",))


def test_header_symbols_known_to_binding():
    from kzg_batch_verification_scheme_b200.api import KzgLib
    assert set(_declared_symbols()) == set(KzgLib.EXPORTS)


def test_product_library_builds_and_exports_everything():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(str(ROOT / "kzg_batch_verification_scheme_b200" / "csrc" / "libkzgb200.so"))
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.kzgb_version.restype = ctypes.c_char_p
    assert b"cuda" in lib.kzgb_version()


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device context creation reports KZGB_ERROR."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kzg_batch_verification_scheme_b200.api import KzgError, load
    lib = load()
    with pytest.raises(KzgError):
        lib.context()


def test_oracle_exports_everything(oracle_lib):
    for name in _declared_symbols():
        assert hasattr(oracle_lib.lib, name), name


def test_package_does_not_reference_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = ROOT / "kzg_batch_verification_scheme_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + list(pkg.rglob("Makefile")):
        txt = p.read_text()
        for line in txt.splitlines():
            code = line.split("//")[0].split("#")[0] if p.suffix != ".py" else line.split("#")[0]
            assert "oracle/" not in code and "libkzgb_oracle" not in code and "import oracle" not in code, (p, line)
