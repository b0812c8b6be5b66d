#!/usr/bin/env python3
"""G1 MSM rate (BASELINE.json metric "G1 MSM Mpts/s") at m = 2^16 .. 2^24 points, 255- and 128-bit scalars, through
kzgb_g1_msm (host buffers; the reported times are the device stage times of the call: digits + sort, accumulate,
reduce + combine).  Points: decompressed commitments of the synthetic stream (valid subgroup points)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from kzg_batch_verification_scheme_b200.api import load  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24]
lib = load()
mmax = 1 << max(sizes)
ctx = lib.test_context(n_max=mmax)
peak, _ = ctx.imad_peak()
C, Z, Y, PI = ctx.synth_instance(0x4B5A4706, 0, mmax)
rc, aff, st = ctx.g1_decompress_batch(C)
assert rc == 0 and not any(st)
del C, Z, Y, PI
rng = np.random.default_rng(7)
out = {}
for lg in sizes:
    m = 1 << lg
    for nbits in (255, 128):
        sc = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
        if nbits == 255:
            sc[:, 0] &= 0x3F
        else:
            sc[:, :16] = 0
        scb = sc.tobytes()
        best = None
        for _ in range(3):
            rc, res = ctx.g1_msm(aff[:96 * m], scb, nbits)
            assert rc == 0
            t = ctx.g1_msm_times()
            best = t if best is None or t[3] < best[3] else best
        ipp = 52.2e3 if nbits == 255 else 29.4e3
        row = {"mpts_per_s": m / (best[3] * 1e-3) / 1e6, "ms": {"sort": best[0], "accumulate": best[1], "reduce": best[2], "total": best[3]},
               "imad_frac_survey_model_c16": m * ipp / (best[3] * 1e-3) / 9.27e12}
        out[f"2^{lg}_{nbits}bit"] = row
        print(f"m=2^{lg} {nbits}-bit: {row['mpts_per_s']:.1f} Mpts/s  {json.dumps(row['ms'])}", flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "msm_probe.json").write_text(json.dumps(out, indent=1))
