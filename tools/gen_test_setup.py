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
# extended test setup for the cell batch (BASELINE.json config[4]): [tau^j]G1 j < 64 | [tau^j]G2 j <= 64
ext = k.setup_g1(64) + k.setup_g2(65)
assert len(ext) == 64 * 48 + 65 * 96 and ext[:48] == blob[:48] and ext[64 * 48:64 * 48 + 192] == blob[48:]
p = ROOT / "kzg_batch_verification_scheme_b200" / "data" / "test_setup_cells.bin"
p.write_bytes(ext)
print("wrote", p)
