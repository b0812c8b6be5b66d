"""The C-ABI libraries load and export every symbol include/kzgb200.h declares (no compute without a GPU)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "kzgb200.h").read_text()
    names = set(re.findall(r"\b(kzgb_[a-z0-9_]+|verify_[a-z_]*kzg_proof[a-z0-9_]*)\s*\(", txt))
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
        lib.test_context()


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


def test_product_host_root_hash_matches_spec():
    """kzgb_fs_root of the CUDA library is host code (SHA extensions when the CPU has them, portable rounds
    otherwise): both paths against hashlib, no GPU needed."""
    import hashlib
    import os
    import random
    import subprocess
    import sys
    from kzg_batch_verification_scheme_b200.api import load
    lib = load()

    def fs_root(dig, n_total):
        out = ctypes.create_string_buffer(32)
        assert lib.lib.kzgb_fs_root(out, dig, len(dig) // 32, n_total) == 0
        return out.raw
    rnd = random.Random(9)
    cases = [(k, bytes(rnd.randrange(256) for _ in range(32 * k))) for k in (1, 2, 3, 7, 64, 1000, 8191)]
    for k, dig in cases:
        n_total = k * 128 - 5
        want = hashlib.sha256(b"KZGB200/root_v1_" + (4096).to_bytes(8, "big") + n_total.to_bytes(8, "big") + dig).digest()
        assert fs_root(dig, n_total) == want, k
    # the portable path in a fresh process
    code = ("import ctypes,hashlib,sys; sys.path.insert(0, %r); from kzg_batch_verification_scheme_b200.api import load; l = load(); "
            "d = bytes(range(256)) * 13; n = 104 * 128; "
            "w = hashlib.sha256(b'KZGB200/root_v1_' + (4096).to_bytes(8, 'big') + n.to_bytes(8, 'big') + d).digest(); "
            "o = ctypes.create_string_buffer(32); assert l.lib.kzgb_fs_root(o, d, len(d) // 32, n) == 0; "
            "assert o.raw == w; print('ok')") % str(ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, KZGB_NO_SHANI="1"), capture_output=True, text=True)
    assert out.stdout.strip() == "ok", out.stderr
