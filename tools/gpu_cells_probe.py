#!/usr/bin/env python3
"""Timing of the cell batch (config[4]: 128 blobs x 128 cells) and of the blob batch on one GPU; prints wall and
device stage times.  Inputs come from the oracle's generators (test infrastructure)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from kzg_batch_verification_scheme_b200.api import KzgLib, load  # noqa: E402
from tests.test_cells_oracle import synth_cells  # noqa: E402

olib = KzgLib(ROOT / "oracle" / "libkzgb_oracle.so")
g1, g2 = olib.synth_setup(64, 65)
ctx = load().context(g1, g2, n_max=1 << 15)
comms, ci, xi, cells, proofs = synth_cells(olib, 0x4B5A4704, 128, 128, 4096)
for it in range(4):
    t0 = time.perf_counter()
    r = ctx.verify_cell_kzg_proof_batch(comms, ci, xi, cells, proofs)
    dt = (time.perf_counter() - t0) * 1e3
    a = ctx.last_artifacts()
    print("cells", r, f"wall {dt:.2f} ms", {k: round(v, 2) for k, v in a["stage_ms"].items() if v})
