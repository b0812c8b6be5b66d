#!/usr/bin/env python3
"""Quick on-box measurement: IMAD peak, stage times at a few n (device-resident inputs), MSM rate."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from kzg_batch_verification_scheme_b200.api import load  # noqa: E402

out = {}
lib = load()
sizes = [int(a) for a in sys.argv[1:]] or [4096, 1 << 16]
ctx = lib.test_context(n_max=max(sizes))
peak, ms = ctx.imad_peak()
out["imad_peak_per_s"] = peak
print(f"IMAD.WIDE peak: {peak / 1e12:.2f} T/s  ({ms:.2f} ms)", flush=True)
for n in sizes:
    bufs = [torch.empty(s * n, dtype=torch.uint8, device="cuda") for s in (48, 32, 32, 48)]
    t0 = time.time()
    ctx.synth_instance(0x4B5A4701, 0, n, device_ptrs=tuple(t.data_ptr() for t in bufs))
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    ptrs = [t.data_ptr() for t in bufs]
    stream = torch.cuda.current_stream().cuda_stream
    best = None
    for it in range(4):
        t0 = time.time()
        rc, ok = ctx.verify_kzg_proof_batch_device(*ptrs, n, stream)
        wall = (time.time() - t0) * 1e3
        st = ctx.last_stage_ms()
        if best is None or st["total"] < best["total"]:
            best = dict(st, wall_ms=wall, rc=rc, ok=ok)
    best["gen_s"] = gen_s
    best["proofs_per_s"] = n / (best["total"] * 1e-3)
    out[f"n={n}"] = best
    print(n, json.dumps(best), flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "quick.json").write_text(json.dumps(out, indent=1))
