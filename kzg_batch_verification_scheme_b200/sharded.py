"""One-process-per-GPU driver of the shard-level C ABI (BASELINE.json:5 multi-GPU: contiguous shards, host
combine of 320-byte partials, no NCCL on the data path).  The only exchanges are tiny host byte strings
(chunk digests, partials) moved with torch.distributed object collectives over the CPU (gloo) group.

    rank r owns proofs [r*n_local, (r+1)*n_local) of ONE batch of world*n_local proofs
    phase 1 (local)   : K1 + leaf/chunk hashes                  -> 32 B digest per 128 proofs
    all_gather        : digests  (256 KiB per 2^20 proofs)
    root (every rank) : SHA-256 over all digests (host)
    phase 2 (local)   : challenges, three MSMs                  -> 320-byte partial
    gather to rank 0  : partials
    rank 0            : combine + two-pairing check             -> verdict
"""
from __future__ import annotations


def sharded_verify(ctx, dist, rank: int, world: int, C, z, y, pi, n_local: int, on_device: bool = False, stream: int = 0,
                   slot: int = 0):
    """Returns (rc, ok) on rank 0 and (rc, None) elsewhere.  `ctx` is an api.Context of any library that
    exports include/kzgb200.h; `dist` is torch.distributed (initialised) or None when world == 1."""
    n_total = n_local * world
    rc, dig, _ = ctx.shard_phase1(slot, C, z, y, pi, n_local, on_device=on_device, stream=stream)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (rc, dig))
        rc = max(g[0] for g in gathered)
        digs = b"".join(g[1] for g in gathered)
    else:
        digs = dig
    if rc:
        return rc, (False if rank == 0 else None)
    root = ctx.fs_root(digs, n_total)
    rc, part = ctx.shard_phase2(slot, root, rank * n_local)
    if world > 1:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((rc, part), parts, dst=0)
        if rank != 0:
            return rc, None
        rc = max(p[0] for p in parts)
        blob = b"".join(p[1] for p in parts)
    else:
        blob = part
    if rc:
        return rc, False
    return ctx.combine_verify(blob)
