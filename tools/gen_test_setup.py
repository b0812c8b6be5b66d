#!/usr/bin/env python3
"""Writes the insecure TEST trusted setup (known tau) used by tests and bench:
kzg_batch_verification_scheme_b200/data/test_setup.bin = [tau^0]G1 (48 B) | [tau^0]G2 | [tau^1]G2 (96 B each),
and the same bytes under tests/golden/.  Generated with the pure-Python model (oracle/pymodel)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pymodel import kzg_model as k  # noqa: E402

blob = k.setup_g1(1) + k.setup_g2(2)
assert len(blob) == 240
for p in (ROOT / "kzg_batch_verification_scheme_b200" / "data" / "test_setup.bin", ROOT / "tests" / "golden" / "test_setup.bin"):
    p.write_bytes(blob)
    print("wrote", p)
