#!/usr/bin/env python3
"""bench.py -- verified KZG proofs/s (BASELINE.json metric) on 1..8 B200.

A "step" is one batch verification (decompress + subgroup checks, Fiat-Shamir, three MSMs, pairing check) through
the C ABI of libkzgb200.so.  `value` = device-resident inputs; `e2e` = plain (pageable) host buffers through
verify_kzg_proof_batch, H2D inside the timed region (`e2e_pinned`: the same from pinned buffers).  N>1 (torchrun,
one process per GPU): default STRONG scaling, BASELINE.json config[3] -- ONE batch of 2^n proofs cut into N
contiguous shards; chunk digests, 66 pairing terms per shard and the verdicts cross the host through a
shared-memory mailbox (no NCCL; gloo only for barriers and the timing scalar); rank 0 adds the terms and runs the
pairing check.  The weak figure (2^n proofs per GPU in one batch) is measured as a second pass of the same line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n LOG2] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# algorithmic work model (DESIGN.md "Work model").  Unit = one 32x32->64 multiply-add.
# Generic Montgomery product: 2N^2+N = 300 (N = 12); dedicated squaring: N(N-1)/2 + N + N^2 + N = 234.
# K1 per point: sqrt a^((p+1)/4) = 377 S + 85 M (sliding 4-bit windows), on-curve check 2 S + 1 M, two |x| chains =
# 126 doublings (2M+5S) + 5 mixed additions (8M+3S) + 5 full additions (12M+4S) + compare (3M+1S):
IMAD_PER_FPMUL, IMAD_PER_FPSQR = 300, 234
K1_M_PER_POINT = 85 + 1 + 126 * 2 + 5 * 8 + 5 * 12 + 3 + 2           # + to/from Montgomery
K1_S_PER_POINT = 377 + 2 + 126 * 5 + 5 * 3 + 5 * 4 + 1
K1_IMAD_PER_POINT = K1_M_PER_POINT * IMAD_PER_FPMUL + K1_S_PER_POINT * IMAD_PER_FPSQR
# batches of >= KZGB_SG_BATCH_MIN (default 2) proofs: the subgroup check is done on 128 bucket-slice sums per MSM,
# K1 is the decompression kernel alone (sqrt + on-curve + Montgomery conversions)
K1A_M_PER_POINT, K1A_S_PER_POINT = 85 + 1 + 2, 377 + 2
K1A_IMAD_PER_POINT = K1A_M_PER_POINT * IMAD_PER_FPMUL + K1A_S_PER_POINT * IMAD_PER_FPSQR
K1_IMAD_PER_POINT_SURVEY = (471 + 1021) * 300                          # SURVEY.md 8(d) model, M = S = 300
MSM_FPMUL_PER_PROOF = 370                                              # SURVEY.md App. C, n = 2^20
SEED = 0x4B5A4703


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=20, help="log2 of proofs per GPU (default 20)")
    ap.add_argument("--impl", default="kzgb200", choices=["kzgb200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = one batch of 2^n proofs cut N ways (BASELINE.json config[3], default); weak = 2^n per GPU. "
                         "The other mode is measured as a second pass of the same line.")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / msm / n=2^16 extras")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return None
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_lib():
    """CPU oracle: ONLY for the cpu_baseline leg and --impl reference (test infrastructure otherwise)."""
    from kzg_batch_verification_scheme_b200.api import KzgLib
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    return KzgLib(ROOT / "oracle" / "libkzgb_oracle.so")


def time_oracle(octx, n_sample, threads, reps=1):
    """proofs/s of the CPU oracle on a sample of the bench workload (same generator stream)."""
    octx.set_threads(threads)
    C, Z, Y, PI = octx.synth_instance(SEED, 0, n_sample)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, ok = octx.verify_kzg_proof_batch(C, Z, Y, PI, n_sample)
        dt = time.perf_counter() - t0
        assert (rc, ok) == (0, True), (rc, ok)
        best = dt if best is None else min(best, dt)
    return n_sample / best, best


def workload_name(n_log2, world, scaling, n_total):
    """The `config.workload` string: the same for the product arm and the reference arm."""
    multi = world > 1
    return (f"BLS12-381 KZG batch verify, ONE batch of n=2^{n_log2} proofs" +
            (f" cut into {world} contiguous shards (BASELINE.json config[3])" if multi and scaling == "strong" else
             f" per GPU ({n_total} in one batch)" if multi else "") +
            ", compressed inputs incl. decompression + subgroup checks, Fiat-Shamir, 3 MSMs, pairing check")


def run_reference(args):
    """Reference arm: the upstream reference has no implementation (LICENSE only), so the CPU oracle port
    is what is timed, on all host threads, on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = oracle_lib()
    octx = lib.test_context()
    cores = os.cpu_count() or 1
    # calibrate a sample that takes a few seconds per step
    rate, _ = time_oracle(octx, 1024, cores)
    n_sample = int(min(1 << args.n, max(1024, 2 ** (int(rate * 4).bit_length() - 1)))) if rate >= 1 else 1024
    octx.set_threads(cores)
    C, Z, Y, PI = octx.synth_instance(SEED, 0, n_sample)              # inputs resident in host memory before timing
    for _ in range(min(args.warmup, 1)):
        assert octx.verify_kzg_proof_batch(C, Z, Y, PI, n_sample) == (0, True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert octx.verify_kzg_proof_batch(C, Z, Y, PI, n_sample) == (0, True)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_sample / dt
    line = {
        "impl": "reference", "metric": "verified KZG proofs/s", "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.gpus, args.scaling, (1 << args.n) * (args.gpus if args.scaling == "weak" else 1)),
                   "sample": f"CPU arm: each step = one batch of {n_sample} proofs of the same generator stream (a 2^{args.n} batch takes ~20 s on 16 "
                             "cores); proofs/s is size-independent to within the window-width effect"},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
                         "sample": f"one batch of {n_sample} proofs of the bench stream per step, {args.steps} steps, all host threads"},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from kzg_batch_verification_scheme_b200.api import CHUNK, PARTIAL_BYTES, TERMS_BYTES, load
    from kzg_batch_verification_scheme_b200.sharded import HostMailbox, sharded_verify

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    torch.cuda.set_device(local)
    if multi:
        # host-side plumbing only (barriers, the timing scalar, the name of the shared-memory mailbox): no NCCL anywhere
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    n_cfg = 1 << args.n
    lib = load()
    ctx = lib.test_context(devices=[local], n_max=n_cfg)
    box = HostMailbox(dist, rank, world, n_cfg) if multi else None
    stream = torch.cuda.current_stream().cuda_stream

    class Workload:
        """Shard of this rank: device-resident bytes, a pinned and a plain (pageable) host mirror."""
        def __init__(self, n_local):
            self.n_local, self.n_total = n_local, n_local * world
            self.dbuf = [torch.empty(s * n_local, dtype=torch.uint8, device="cuda") for s in (48, 32, 32, 48)]
            ctx.synth_instance(SEED, rank * n_local, n_local, device_ptrs=tuple(t.data_ptr() for t in self.dbuf))
            torch.cuda.synchronize()
            self.pinned = [torch.empty(s * n_local, dtype=torch.uint8).pin_memory() for s in (48, 32, 32, 48)]
            self.plain = [torch.empty(s * n_local, dtype=torch.uint8) for s in (48, 32, 32, 48)]       # malloc'ed, pageable
            for h, d in zip(self.pinned, self.dbuf):
                h.copy_(d)
            torch.cuda.synchronize()
            for h, q in zip(self.plain, self.pinned):
                h.copy_(q)
            self.dptr = [t.data_ptr() for t in self.dbuf]
            self.pptr = [t.data_ptr() for t in self.pinned]
            self.hptr = [t.data_ptr() for t in self.plain]

    host_trace = {}

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
            box.barrier()            # shared-memory rendezvous right after: the ranks leave within microseconds of each other
        torch.cuda.synchronize()

    def step(w, ptrs, on_device):
        """one batch verification; returns the verdict (the same on every rank)"""
        if not multi:
            if on_device:
                rc, ok = ctx.verify_kzg_proof_batch_device(*ptrs, w.n_local, stream)
            else:
                rc, ok = ctx.verify_kzg_proof_batch(*ptrs, w.n_local)
        else:
            rc, ok = sharded_verify(ctx, dist, rank, world, *ptrs, w.n_local, on_device=on_device, stream=stream, box=box, trace=host_trace)
        assert rc == 0, rc
        return ok

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # 2 x the 126 MB L2

    def timed(w, ptrs, on_device, steps, warmup):
        """K steps, each bracketed by barrier + synchronize and timed with CUDA events; between steps the L2 is flushed
        (untimed) by writing a 256 MB buffer, so no step finds its inputs or tables in cache."""
        for _ in range(warmup):
            assert step(w, ptrs, on_device)
        stage_acc = {}
        ms, launches = 0.0, 0
        for _ in range(steps):
            flush_buf.fill_(1)
            barrier()
            l0 = ctx.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ok = step(w, ptrs, on_device)
            e1.record()
            barrier()
            assert ok, "batch must verify"
            ms += e0.elapsed_time(e1)
            launches += ctx.launch_count() - l0
            for k, v in ctx.last_stage_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        if multi:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            t = torch.tensor([launches], dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            launches = int(t.item())
        return ms / steps, {k: v / steps for k, v in stage_acc.items()}, launches

    def planted(w):
        """a corrupted proof (on the LAST rank for N > 1) must be rejected -- on every rank"""
        if rank == world - 1:
            saved = w.dbuf[3][:48].clone()
            w.dbuf[3][:48] = w.dbuf[3][48:96]
            torch.cuda.synchronize()
        if multi:
            rc, ok = sharded_verify(ctx, dist, rank, world, *w.dptr, w.n_local, on_device=True, stream=stream, box=box)
        else:
            rc, ok = ctx.verify_kzg_proof_batch_device(*w.dptr, w.n_local, stream)
        if rank == world - 1:
            w.dbuf[3][:48] = saved
            torch.cuda.synchronize()
        good = (rc == 0 and not ok)
        if multi:
            t = torch.tensor([int(good)], dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            good = bool(t.item())
        return good

    # ---- the headline pass.  strong (BASELINE.json config[3]): ONE batch of 2^n proofs cut world ways;
    # weak: every rank owns 2^n proofs of one batch of world * 2^n
    n_local = n_cfg if args.scaling == "weak" else max(CHUNK, n_cfg // world)
    w = Workload(n_local)
    n_total = w.n_total
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, stages, launches = timed(w, w.dptr, True, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    host_phase_ms = {k: v / (args.steps + args.warmup) for k, v in host_trace.items()}     # rank 0's host wall time per phase
    ms_e2e, stages_e2e, _ = timed(w, w.hptr, False, args.steps, max(1, args.warmup // 2))
    ms_pin, stages_pin, _ = timed(w, w.pptr, False, args.steps, max(1, args.warmup // 2))
    reject_ok = planted(w)
    other = None
    if multi:
        # the other scaling mode as a second timed pass of the same line
        n_other = n_cfg if args.scaling == "strong" else max(CHUNK, n_cfg // world)
        w.pinned = w.plain = None                            # free the host mirrors of the first pass
        w2 = Workload(n_other)
        ms2, st2, _ = timed(w2, w2.dptr, True, args.steps, args.warmup)
        ms2e, _, _ = timed(w2, w2.hptr, False, args.steps, 1)
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "n_per_gpu": n_other, "n_total": w2.n_total,
                 "value": w2.n_total / (ms2 * 1e-3), "unit": "proofs/s", "ms_per_step": ms2, "stage_ms": st2,
                 "e2e": {"value": w2.n_total / (ms2e * 1e-3), "ms_per_step": ms2e}, "planted_invalid_rejected": planted(w2)}
        del w2
        # the same 2^n batch through the IN-PROCESS multi-device path of the C ABI: ONE context over all N devices, one call
        # of verify_kzg_proof_batch on pinned host buffers (one host thread per device inside the library, terms combined
        # on the host).  Rank 0 drives it; the other ranks wait.
        inproc = None
        if rank == 0:
            try:
                ictx = lib.test_context(devices=list(range(world)), n_max=max(1 << 15, (n_cfg + world - 1) // world + CHUNK), cells=True)
                full = [torch.empty(s * n_cfg, dtype=torch.uint8).pin_memory() for s in (48, 32, 32, 48)]
                tmp = [torch.empty(s * n_cfg, dtype=torch.uint8, device="cuda") for s in (48, 32, 32, 48)]
                ctx.synth_instance(SEED, 0, n_cfg, device_ptrs=tuple(t.data_ptr() for t in tmp))
                torch.cuda.synchronize()
                for h, d in zip(full, tmp):
                    h.copy_(d)
                del tmp
                fptr = [t.data_ptr() for t in full]
                for _ in range(args.warmup):
                    assert ictx.verify_kzg_proof_batch(*fptr, n_cfg) == (0, True)
                times = []
                for _ in range(args.steps):
                    t0 = time.perf_counter()
                    assert ictx.verify_kzg_proof_batch(*fptr, n_cfg) == (0, True)
                    times.append((time.perf_counter() - t0) * 1e3)
                # config[4] "on 8xB200": the 2^14 openings of the cell batch sharded over the same devices (random evaluations
                # with valid commitments and proofs: the full work, verdict "false"), plain host buffers
                import numpy as np
                rngc = np.random.default_rng(11)
                m_ = 128 * 128
                cells_t = torch.from_numpy(rngc.integers(0, 256, size=(m_, 64, 32), dtype=np.uint8))
                cells_t[:, :, 0] &= 0x3F
                cells_t = cells_t.contiguous()
                comm_b, proof_b = bytes(full[0][:48 * 128].numpy()), bytes(full[3][:48 * m_].numpy())
                ci_, xi_ = [k // 128 for k in range(m_)], [k % 128 for k in range(m_)]
                best_c = None
                for _ in range(5):
                    rc, okc = ictx.verify_cell_kzg_proof_batch(comm_b, ci_, xi_, cells_t.data_ptr(), proof_b)
                    assert rc == 0 and okc is False
                    dms = ictx.last_stage_ms()["total"]
                    best_c = dms if best_c is None or dms < best_c else best_c
                cell_multi = {"openings": m_, "devices": world, "device_ms": best_c, "openings_per_s": m_ / (best_c * 1e-3),
                              "host_memory": "pageable"}
                full[3][:48] = full[3][48:96]
                rej = ictx.verify_kzg_proof_batch(*fptr, n_cfg) == (0, False)
                ms_i = sum(times) / len(times)
                inproc = {"n_total": n_cfg, "devices": world, "ms_per_step": ms_i, "value": n_cfg / (ms_i * 1e-3), "unit": "proofs/s",
                          "host_memory": "pinned", "timing": "host wall clock around the blocking call (H2D on every device inside)",
                          "planted_invalid_rejected": rej, "cell_batch_128x128": cell_multi}
                ictx.close()
            except Exception as ex:                                  # noqa: BLE001
                inproc = {"error": repr(ex)}
        barrier()
        if other is not None:
            other_in = inproc
    dbuf, hbuf, dptr = w.dbuf, w.pinned, w.dptr

    value = n_total / (ms_dev * 1e-3)
    e2e_value = n_total / (ms_e2e * 1e-3)
    nch = (n_local + CHUNK - 1) // CHUNK

    extras, configs = {}, {}
    roofline = roofline_hbm = roofline_hbm_k6 = cpu_baseline = None
    msm_top = {}
    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        peaks = json.loads(peaks_path.read_text()) if peaks_path.exists() else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        imad_peak, imad_ms = ctx.imad_peak(wide=True)          # IMAD.WIDE.U32.X carry chains (what the kernels issue)
        imad32_peak, _ = ctx.imad_peak(wide=False)             # plain 32-bit IMAD, context only
        peak_src = "kzgb_imad_peak (library microbenchmark)"
        # the stand-alone microbenchmark keeps each carry chain contiguous in the schedule and reaches a
        # slightly higher rate than the in-library one; a peak is a maximum, so take the larger of the two
        mb = ROOT / "tools" / "microbench" / "imad_flavours"
        if mb.exists():
            try:
                for ln in subprocess.run([str(mb)], capture_output=True, text=True, timeout=60).stdout.splitlines():
                    if "IMAD.WIDE.X" in ln:
                        v = float(ln.split("ms")[1].split("Tops/s")[0]) * 1e12
                        if v > imad_peak:
                            imad_peak, peak_src = v, "tools/microbench/imad_flavours (stand-alone, same run)"
            except Exception:                                   # noqa: BLE001
                pass
        k1_ms = stages.get("decompress", 0.0)
        sg_min = int(os.environ.get("KZGB_SG_BATCH_MIN", "2"))
        sg_batch = sg_min > 0 and n_local >= sg_min
        if k1_ms > 0:
            per_point = K1A_IMAD_PER_POINT if sg_batch else K1_IMAD_PER_POINT
            k1_imad = 2 * n_local * per_point
            ach = k1_imad / (k1_ms * 1e-3)
            kernel = ("K1 = k_decompress_sqrt (one launch over commitments and proofs; the stage also carries the side-stream hashes, challenges "
                      "and sorts that run under it); subgroup membership is established on 128 bucket-slice sums per MSM (batched check), "
                      "not per point") if sg_batch else \
                "K1 = k_decompress_sqrt + k_subgroup_chain1 + k_subgroup_chain2 (one stage, 3 launches)"
            roofline = {"bound": "imad", "kernel": kernel,
                        "achieved": ach / 1e12, "peak": imad_peak / 1e12, "unit": "T wide-IMAD/s", "frac": ach / imad_peak,
                        # dram__bytes_read.sum + dram__bytes_write.sum of k_decompress_sqrt from the ncu --set full capture of THIS build
                        # at the benchmarked n = 2^20 (profiles/r2_k1a_decompress_sqrt_ncu_full_n1048576.txt: 103.7 + 186.1 MB for
                        # 2^21 points = 138.2 B/point; algorithmic 48 + 96 + 1 = 145 B/point), scaled by n.  Per-point chains:
                        # r1 capture at n = 65536, 507 B/point
                        "traffic": int(2 * n_local * (138.2 if sg_batch else 507)),
                        "traffic_source": "ncu capture at n = 2^20 committed under profiles/ (bench.py cannot run ncu inside the timed run); scaled by n",
                        "peak_source": peak_src + ": carry-chained mad.lo.cc/madc.hi.cc (SASS IMAD.WIDE.U32.X) on all SMs, measured in this "
                                       "run; 32 lanes/clk/SM on B200 (148 x 32 x 1.965 GHz = 9.31 T/s nominal)",
                        "imad32_issue_peak": imad32_peak / 1e12,
                        "frac_vs_imad32_issue_peak": ach / imad32_peak,      # SURVEY.md 8(d)'s a-priori denominator (64 plain IMAD/clk/SM)
                        "kernel_alone_ms_ncu": 27.61 if (sg_batch and n_local == 1 << 20) else None,   # same capture: the kernel without the side streams
                        "algorithmic_per_launch": k1_imad, "launch_ms": k1_ms,
                        "formula": (f"2n x ({K1A_M_PER_POINT} M x 300 + {K1A_S_PER_POINT} S x 234)" if sg_batch else
                                    f"2n x ({K1_M_PER_POINT} M x 300 + {K1_S_PER_POINT} S x 234)") +
                                   " wide multiply-adds; stage time from CUDA events on the library's stream",
                        "subgroup_check": "batched (bucket slices)" if sg_batch else "per point",
                        "whole_batch_frac": (n_local * (2 * per_point + MSM_FPMUL_PER_PROOF * 300)) / (ms_dev * 1e-3) / imad_peak,
                        # the same batch against SURVEY.md 8(d)'s work model (per-point subgroup chains: 2 x 447.6 k + 111 k
                        # multiply-adds per proof): > 1 means the batched subgroup check beats the model's own "100 %" bound
                        "whole_batch_frac_survey_model": (n_local * (2 * K1_IMAD_PER_POINT_SURVEY + MSM_FPMUL_PER_PROOF * 300)) / (ms_dev * 1e-3) / imad_peak}
            if not sg_batch:
                roofline["frac_survey_model"] = 2 * n_local * K1_IMAD_PER_POINT_SURVEY / (k1_ms * 1e-3) / imad_peak
            k1_bytes = 2 * n_local * ((48 + 96 + 1) if sg_batch else (48 + 96 + 1 + 2 * 144 + 2 * 96))
            roofline_hbm = {"bound": "hbm", "kernel": "K1", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                            "note": "K1 is integer-pipe bound by construction (hundreds of Fp products per ~150-625 bytes moved); HBM fraction reported for completeness"}
            # K6 point stream (BJ:5 "achieved HBM GB/s for the point stream"): every sorted entry reads its key, value and one 96-byte
            # affine point; 4 n W entries over the three sums (S2' has 2(n+1) points).  Measured DRAM traffic of pass 1 is 1.5x
            # this (points on 16-byte alignment straddle 32-byte sectors: profiles/r2_k6_accumulate_pass1_ncu_full_n1048576.txt)
            acc_ms = stages.get("msm_accumulate", 0.0)
            if acc_ms > 0:
                from math import ceil, log2
                c_w = max(3, min(16, int(log2(max(n_local, 2))) - 3))
                entries = (2 * n_local + 2 * (n_local + 1)) * ceil(128 / c_w)
                k6_bytes = entries * (96 + 8)
                roofline_hbm_k6 = {"bound": "hbm", "kernel": "K6 accumulate (pass 1 + pass 2 + slice sums of the three sums)", "achieved": k6_bytes / (acc_ms * 1e-3) / 1e9,
                                   "peak": hbm_peak, "unit": "GB/s", "frac": k6_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak,
                                   "algorithmic_bytes": k6_bytes, "stage_ms": acc_ms,
                                   "note": "integer-pipe bound (2868 multiply-adds per 104 bytes); reported because BASELINE.json:5 asks for it"}
        if not args.no_extras:
            # CPU baseline: oracle port on the box's host cores, bounded sample (~10-20 s)
            try:
                olib = oracle_lib()
                octx = olib.test_context()
                cores = os.cpu_count() or 1
                rate, _ = time_oracle(octx, 1024, cores)
                n_s = 1024
                while n_s * 2 <= n_local and n_s * 2 / rate < 12.0:
                    n_s *= 2
                rate, secs = time_oracle(octx, n_s, cores)
                cpu_baseline = {"value": rate, "unit": "proofs/s", "cores": cores, "kind": "port",
                                "sample": f"one batch of {n_s} proofs (first {n_s} of the bench stream), {secs:.1f} s, all host threads"}
                r1, s1 = time_oracle(octx, max(256, n_s // max(cores, 1) // 2), 1)
                cpu_baseline["single_thread_value"] = r1
                octx.close()
            except Exception as ex:                              # noqa: BLE001
                cpu_baseline = {"error": repr(ex)}
            # the other single-GPU configurations BASELINE.json names, one batch at a time, device-resident inputs
            # (best of 5, L2 flushed before each): config[1] n = 4096, config[2] n = 2^16 with its planted invalid proof
            if not multi:
                for lg in (12, 16):
                    if lg >= args.n:
                        continue
                    n2 = 1 << lg
                    best = None
                    for _ in range(5):
                        flush_buf.fill_(1)
                        torch.cuda.synchronize()
                        rc, ok = ctx.verify_kzg_proof_batch_device(*dptr, n2, stream)
                        assert (rc, ok) == (0, True)
                        st_ = ctx.last_stage_ms()
                        best = st_ if best is None or st_["total"] < best["total"] else best
                    entry = {"ms": best["total"], "proofs_per_s": n2 / (best["total"] * 1e-3), "stage_ms": best}
                    if lg == 16:
                        saved = dbuf[3][48 * 777:48 * 778].clone()
                        dbuf[3][48 * 777:48 * 778] = dbuf[3][48 * 778:48 * 779]
                        rc, ok = ctx.verify_kzg_proof_batch_device(*dptr, n2, stream)
                        entry["planted_invalid_rejected"] = (rc == 0 and not ok)
                        dbuf[3][48 * 777:48 * 778] = saved
                        torch.cuda.synchronize()
                    configs[f"n{n2}"] = entry
                # config[4]: PeerDAS-shaped cell batch, 128 blobs x 128 cells = 2^14 openings, plain host buffers.  Valid
                # commitments and proofs of the bench stream with random evaluations: the full work, verdict "false"
                import numpy as np
                cctx = lib.test_context(devices=[local], n_max=1 << 15, cells=True)
                rngc = np.random.default_rng(11)
                nb_, nc_ = 128, 128
                m_ = nb_ * nc_
                cells_t = torch.from_numpy(rngc.integers(0, 256, size=(m_, 64, 32), dtype=np.uint8))
                cells_t[:, :, 0] &= 0x3F
                cells_t = cells_t.contiguous()
                comm_b = bytes(torch.empty(48 * nb_, dtype=torch.uint8).copy_(dbuf[0][:48 * nb_]).numpy())
                proof_b = bytes(torch.empty(48 * m_, dtype=torch.uint8).copy_(dbuf[3][:48 * m_]).numpy())
                ci_ = [k // nc_ for k in range(m_)]
                xi_ = [k % nc_ for k in range(m_)]
                best_w = best_d = None
                for _ in range(5):
                    flush_buf.fill_(1)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    rc, okc = cctx.verify_cell_kzg_proof_batch(comm_b, ci_, xi_, cells_t.data_ptr(), proof_b)
                    dt = (time.perf_counter() - t0) * 1e3
                    assert rc == 0 and okc is False
                    dms = cctx.last_stage_ms()["total"]
                    best_w = dt if best_w is None or dt < best_w else best_w
                    best_d = dms if best_d is None or dms < best_d else best_d
                configs["cell_batch_128x128"] = {"openings": m_, "device_ms": best_d, "wall_ms_incl_binding": best_w,
                                                 "openings_per_s": m_ / (best_d * 1e-3), "h2d_bytes": 2048 * m_ + 48 * (m_ + nb_) + 8 * m_,
                                                 "host_memory": "pageable", "gpus": 1,
                                                 "note": "device_ms: CUDA events around the whole call incl. the 33.5 MB H2D copy; "
                                                         "wall also counts the ctypes marshalling of the index lists"}
                cctx.close()
            if not multi:
                # several batches in flight through the LIBRARY's submit / wait API (one context, one caller thread): the
                # latency-bound tail (bucket reduction, pairing) of batch k runs under K1 of batch k+1
                pipe = {}
                for lg, nb in ((12, 96), (16, 36), (args.n, 8)):
                    if lg > args.n:
                        continue
                    npl = 1 << lg
                    # small batches are bound by the one-SM serial pairing kernel (0.7 ms): more of them in flight run those
                    # kernels side by side on different SMs
                    depth = 8 if lg <= 12 else 3 if lg <= 16 else 2
                    assert ctx.pipeline_init(depth) == 0

                    def run(count):
                        pend = []
                        for _ in range(count):
                            if len(pend) == depth:
                                assert ctx.verify_kzg_proof_batch_wait(pend.pop(0)) == (0, True)
                            rc, t = ctx.verify_kzg_proof_batch_submit(*dptr, npl, on_device=True)
                            assert rc == 0
                            pend.append(t)
                        for t in pend:
                            assert ctx.verify_kzg_proof_batch_wait(t) == (0, True)
                    run(depth)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run(nb)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    pipe[f"n=2^{lg}"] = {"in_flight": depth, "batches": nb, "proofs_per_s": nb * npl / (ms * 1e-3), "ms_per_batch": ms / nb,
                                          "api": "kzgb_pipeline_init + verify_kzg_proof_batch_submit / _wait"}
                assert ctx.pipeline_init(1) == 0
                extras["pipelined"] = pipe
            if not multi:
                import numpy as np
                m = min(n_local, 1 << 20)
                rc, aff, st = ctx.g1_decompress_batch(bytes(hbuf[0][:48 * m].numpy().tobytes()))
                assert rc == 0 and not any(st)
                rng = np.random.default_rng(7)
                for nbits in (255, 128):
                    sc = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
                    if nbits == 255:
                        sc[:, 0] &= 0x3F
                    else:
                        sc[:, :16] = 0
                    best = None
                    for _ in range(3):
                        rc, out = ctx.g1_msm(aff, sc.tobytes(), nbits)
                        assert rc == 0
                        t = ctx.g1_msm_times()
                        best = t if best is None or t[3] < best[3] else best
                    extras[f"msm_{nbits}bit_mpts_per_s"] = m / (best[3] * 1e-3) / 1e6
                    extras[f"msm_{nbits}bit_ms"] = {"sort": best[0], "accumulate": best[1], "reduce": best[2], "total": best[3]}
                    ipp = 52.2e3 if nbits == 255 else 29.4e3          # SURVEY 8(d) multiply-adds per point at c=16
                    extras[f"msm_{nbits}bit_imad_frac"] = m * ipp / (best[3] * 1e-3) / imad_peak
                    msm_top[f"{nbits}bit_2^20"] = {"mpts_per_s": m / (best[3] * 1e-3) / 1e6, "ms": best[3],
                                                   "imad_frac_survey_model": m * ipp / (best[3] * 1e-3) / imad_peak}
                    # the same at m = 2^16
                    m16 = 1 << 16
                    b16 = None
                    for _ in range(3):
                        rc, out = ctx.g1_msm(aff[:96 * m16], sc[:m16].tobytes(), nbits)
                        assert rc == 0
                        t = ctx.g1_msm_times()
                        b16 = t if b16 is None or t[3] < b16[3] else b16
                    msm_top[f"{nbits}bit_2^16"] = {"mpts_per_s": m16 / (b16[3] * 1e-3) / 1e6, "ms": b16[3]}
                # blob-level caller: 1024 blobs of 128 KiB from pinned host memory.  Random
                # evaluations with unrelated (valid) commitments and proofs: same work, verdict "false"
                mb = 1024
                blob_t = torch.from_numpy(rng.integers(0, 256, size=(mb, 4096, 32), dtype=np.uint8))
                blob_t[:, :, 0] &= 0x3F
                blob_t = blob_t.contiguous().pin_memory()
                Cb = bytes(torch.empty(48 * mb, dtype=torch.uint8).copy_(dbuf[0][:48 * mb]).numpy())
                Pb = bytes(torch.empty(48 * mb, dtype=torch.uint8).copy_(dbuf[3][:48 * mb]).numpy())
                best_b = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    rc, okb = ctx.verify_blob_kzg_proof_batch(blob_t.data_ptr(), Cb, Pb, m=mb)
                    dt = (time.perf_counter() - t0) * 1e3
                    assert rc == 0 and okb is False
                    best_b = dt if best_b is None or dt < best_b else best_b
                extras["blob_batch"] = {"blobs": mb, "ms": best_b, "blobs_per_s": mb / (best_b * 1e-3),
                                        "h2d_gb_per_s": mb * 131072 / (best_b * 1e-3) / 1e9,
                                        "note": "hash (SHA-256 tree) + barycentric evaluation of 4096 points per blob + plain batch"}
        line = {
            "metric": "verified KZG proofs/s", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.n, world, args.scaling, n_total),
                       "n_per_gpu": n_local, "n_total": n_total, "seed": hex(SEED),
                       "l2": "L2 flushed between steps (256 MB write, outside the timed region); every step is timed on its own",
                       "parallelism": (f"contiguous shards x{world}, one process per GPU; digests, pairing terms and verdicts cross the host "
                                       "(shared-memory mailbox), no NCCL") if multi else "single GPU"},
            "clocks": clocks,
            # e2e: PLAIN (pageable) host buffers through verify_kzg_proof_batch -- the library orders its copies and launches so
            # that K1 starts after the first sixteenth of C; e2e_pinned: the same call on cudaMallocHost'ed buffers
            "e2e": {"value": e2e_value, "unit": "proofs/s", "ms_per_step": ms_e2e, "host_memory": "pageable", "stage_ms": stages_e2e,
                    "h2d_bytes_per_step": 160 * n_local * world,
                    "d2h_bytes_per_step": (32 * nch + (TERMS_BYTES if multi else 0) + 16) * world},
            "e2e_pinned": {"value": n_total / (ms_pin * 1e-3), "unit": "proofs/s", "ms_per_step": ms_pin, "host_memory": "pinned"},
            "gpu_launches": launches,
            "stage_ms": stages,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_hbm_k6": roofline_hbm_k6, "cpu_baseline": cpu_baseline,
            "msm": msm_top,
            "planted_invalid_rejected": reject_ok,
            "configs": configs,
            "extras": extras,
        }
        if other:
            line[other["scaling"]] = other
            line["host_phase_ms"] = host_phase_ms
            line["in_process_multi_device"] = other_in
        print(json.dumps(line), flush=True)
    ctx.close()
    if multi:
        dist.barrier()
        box.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
