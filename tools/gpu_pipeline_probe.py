#!/usr/bin/env python3
"""Probe: throughput with D batches in flight on ONE GPU (D contexts, D host threads, shared inputs)."""
import sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from kzg_batch_verification_scheme_b200.api import load

lib = load()
for lg in (16, 20):
    n = 1 << lg
    bufs = [torch.empty(s * n, dtype=torch.uint8, device="cuda") for s in (48, 32, 32, 48)]
    c0 = lib.test_context(n_max=n)
    c0.synth_instance(0x4B5A4701, 0, n, device_ptrs=tuple(t.data_ptr() for t in bufs))
    torch.cuda.synchronize()
    ptrs = [t.data_ptr() for t in bufs]
    for depth in (1, 2, 3):
        ctxs = [c0] + [lib.test_context(n_max=n) for _ in range(depth - 1)]
        steps = 12 if lg == 16 else 6
        def work(ctx, k):
            for _ in range(k):
                rc, ok = ctx.verify_kzg_proof_batch_device(*ptrs, n, 0)
                assert (rc, ok) == (0, True)
        for c in ctxs: work(c, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        th = [threading.Thread(target=work, args=(c, steps)) for c in ctxs]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"n=2^{lg} depth={depth}: {depth*steps*n/(ms*1e-3)/1e6:.3f} M proofs/s  ({ms/(depth*steps):.2f} ms per batch, wall {wall*1e3/(depth*steps):.2f})", flush=True)
        for c in ctxs[1:]: c.close()
    c0.close()
