"""One-process-per-GPU driver of the shard-level C ABI (BASELINE.json:5 multi-GPU: contiguous shards, combined on the
host, no NCCL on the data path).  The only exchanges are small host byte strings:

    rank r owns the proofs [off_r, off_r + n_r) of ONE batch of sum(n_r) proofs (n_r a multiple of 128 except on the last rank)
    phase 1 (local)   : K1 + leaf/chunk hashes                  -> 32 B digest per 128 proofs
    all-gather        : digests  (256 KiB per 2^20 proofs) and the shard sizes
    root (every rank) : SHA-256 over all digests (host)
    phase 2 (local)   : challenges, three MSMs                  -> 66 pairing terms (12.7 KB) or one 320-byte partial
    gather to rank 0  : terms / partials
    rank 0            : sum of the terms + pairing check        -> verdict
    all-gather        : every shard's input-validation result and rank 0's verdict (every rank returns the same answer)

Two carriers for those byte strings: `HostMailbox`, a shared-memory segment (/dev/shm) with one slot and one sequence
flag per rank -- an exchange costs a memcpy and a few microseconds of polling -- and the torch.distributed CPU group
(gloo tensors; used when no mailbox is given).  Neither touches device memory.
"""
from __future__ import annotations

import mmap
import os
import struct
import time

import numpy as np

from .api import CHUNK, PARTIAL_BYTES, TERMS_BYTES

_HDR = struct.Struct("<qqqqqq")          # rc, n_local, bad points, bad scalars, combine rc, verdict


class HostMailbox:
    """All-gather of small byte strings between the ranks of one node through a shared-memory file.

    Layout: for each of the three message kinds and two alternating parities, `world` flags (one cache line each)
    followed by `world` payload slots.  A rank publishes by copying its payload into its slot and then storing the
    sequence number into its flag (x86 keeps the store order); readers poll the flags.  Every verification ends with
    an all-gather, so a slot is never rewritten before all ranks have read it."""
    KINDS = 4            # digests, terms / partials, verdicts, barrier

    def __init__(self, dist, rank: int, world: int, max_n_local: int, tag: str | None = None):
        self.rank, self.world = rank, world
        nch = (max_n_local + CHUNK - 1) // CHUNK
        self.slot = (_HDR.size + max(32 * nch, TERMS_BYTES, PARTIAL_BYTES) + 63) // 64 * 64
        self.region = 64 * world + self.slot * world
        size = self.region * self.KINDS * 2
        names = [None]
        if rank == 0:
            import tempfile
            base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
            names[0] = os.path.join(base, f"kzgb200_{tag or os.environ.get('MASTER_PORT', '0')}_{os.getpid()}")
            with open(names[0], "wb") as f:
                f.truncate(size)
        if world > 1:
            dist.broadcast_object_list(names, src=0)
        self.path = names[0]
        self.fd = os.open(self.path, os.O_RDWR)
        self.mm = mmap.mmap(self.fd, size)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.flags = [[np.frombuffer(self.mm, dtype=np.int64, count=8 * world, offset=self.region * (2 * k + p)).reshape(world, 8)[:, 0]
                       for p in range(2)] for k in range(self.KINDS)]
        self.seq = [0] * self.KINDS
        if world > 1:
            dist.barrier()                                   # everyone has mapped the file before it is unlinked
        if rank == 0:
            os.unlink(self.path)

    def close(self):
        self.flags = None
        self.buf = None
        try:
            self.mm.close()
        except BufferError:
            pass
        os.close(self.fd)

    def _slot_off(self, kind, parity, r):
        return self.region * (2 * kind + parity) + 64 * self.world + self.slot * r

    def post(self, kind: int, payload) -> int:
        """Publish this rank's payload of the next exchange of `kind`; returns its sequence number."""
        self.seq[kind] += 1
        seq = self.seq[kind]
        p = seq & 1
        off = self._slot_off(kind, p, self.rank)
        n = len(payload)
        assert n <= self.slot
        self.mm[off:off + n] = payload
        self.flags[kind][p][self.rank] = seq
        return seq

    def barrier(self):
        """All ranks have arrived (microseconds of skew; the gloo barrier releases ranks tens of microseconds apart)."""
        self.collect(3, self.post(3, b""), 0)

    def collect(self, kind: int, seq: int, nbytes: int, timeout_s: float = 120.0):
        """Wait until every rank has published exchange `seq` of `kind`; returns one memoryview per rank."""
        f = self.flags[kind][seq & 1]
        spins, t0 = 0, None
        while int(f.min()) < seq:
            spins += 1
            if spins & 0x3FFF == 0:
                t0 = t0 or time.monotonic()
                if time.monotonic() - t0 > timeout_s:
                    raise TimeoutError(f"HostMailbox: rank {self.rank} waited {timeout_s}s for exchange {kind}/{seq}: {f.tolist()}")
        mv = memoryview(self.mm)
        return [mv[self._slot_off(kind, seq & 1, r):self._slot_off(kind, seq & 1, r) + nbytes] for r in range(self.world)]


class _GlooBox:
    """Same interface on the torch.distributed CPU group (fixed-size uint8 tensors, no pickling)."""

    def __init__(self, dist, rank, world):
        import torch
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.pending = {}

    def post(self, kind, payload):
        self.pending[kind] = bytes(payload)
        return 0

    def collect(self, kind, seq, nbytes, timeout_s=0.0):
        torch = self.torch
        mine = torch.zeros(nbytes, dtype=torch.uint8)
        data = self.pending.pop(kind)
        mine[:len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        if self.world == 1:
            return [memoryview(mine.numpy())]
        out = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.world)]
        self.dist.all_gather(out, mine)
        return [memoryview(t.numpy()) for t in out]


def sharded_verify(ctx, dist, rank: int, world: int, C, z, y, pi, n_local: int, on_device: bool = False, stream: int = 0,
                   slot: int = 0, box=None, mode: str = "terms", n_max_local: int | None = None, trace: dict | None = None):
    """Returns (rc, ok) -- the same pair on every rank.  `ctx` is an api.Context of any library that exports
    include/kzgb200.h; `dist` is torch.distributed (initialised) or None when world == 1; `box` a HostMailbox
    (default: the gloo group).  mode "terms": Horner-free exchange of pairing terms; "partials": the 320-byte partials
    of BASELINE.json:5.  Shards may have different sizes; all but the last must be multiples of 128 proofs."""
    t_last = [time.perf_counter()]

    def mark(name):                                          # host wall time per phase (ms), accumulated into `trace`
        if trace is not None:
            now = time.perf_counter()
            trace[name] = trace.get(name, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now
    if box is None:
        box = _GlooBox(dist, rank, world)
    nch_max = ((n_max_local or n_local) + CHUNK - 1) // CHUNK
    if not isinstance(box, HostMailbox) and world > 1 and n_max_local is None:
        # tensor all-gather needs one size on every rank: agree on the largest shard first
        import torch
        t = torch.tensor([n_local], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nch_max = (int(t.item()) + CHUNK - 1) // CHUNK
    # ---- phase 1 + digests
    rc1, dig, _ = ctx.shard_phase1(slot, C, z, y, pi, n_local, on_device=on_device, stream=stream)
    mark("phase1")
    seq = box.post(0, _HDR.pack(rc1, n_local, 0, 0, 0, 0) + dig)
    got = box.collect(0, seq, box.slot if isinstance(box, HostMailbox) else _HDR.size + 32 * nch_max)
    mark("digest_exchange")
    hdrs = [_HDR.unpack_from(g, 0) for g in got]
    sizes = [h[1] for h in hdrs]
    rc = max(h[0] for h in hdrs)
    if any(s <= 0 for s in sizes) or any(s % CHUNK for s in sizes[:-1]):
        rc = max(rc, 1)                                      # KZGB_BADARGS: shard boundaries must fall on chunk boundaries
    if rc:
        return rc, False
    n_total = sum(sizes)
    offset = sum(sizes[:rank])
    digs = b"".join(bytes(g[_HDR.size:_HDR.size + 32 * ((s + CHUNK - 1) // CHUNK)]) for g, s in zip(got, sizes))
    root = ctx.fs_root(digs, n_total)
    mark("root")
    # ---- phase 2 + terms / partials to rank 0
    if mode == "terms":
        rc2, rec = ctx.shard_phase2_terms(slot, root, offset)
        rec_bytes = TERMS_BYTES
    else:
        rc2, rec = ctx.shard_phase2(slot, root, offset)
        rec_bytes = PARTIAL_BYTES
    mark("phase2")
    seq = box.post(1, _HDR.pack(rc2, n_local, 0, 0, 0, 0) + rec)
    rc_c, ok = 0, False
    if rank == 0 or not isinstance(box, HostMailbox):
        got = box.collect(1, seq, _HDR.size + rec_bytes)
        mark("terms_exchange")
        if rank == 0 and not any(_HDR.unpack_from(g, 0)[0] for g in got):
            blob = b"".join(bytes(g[_HDR.size:_HDR.size + rec_bytes]) for g in got)
            rc_c, ok = ctx.combine_verify_terms(blob, world) if mode == "terms" else ctx.combine_verify(blob)
            mark("combine_pairing")
    # ---- every shard's input validation + rank 0's verdict
    if mode == "terms" and rc2 == 0:
        rc3, badp, bads = ctx.shard_finish(slot)
    else:
        rc3, badp, bads = rc2, 0, 0
    mark("finish")
    seq = box.post(2, _HDR.pack(max(rc2, rc3), n_local, badp, bads, rc_c, int(ok)))
    fin = [_HDR.unpack_from(g, 0) for g in box.collect(2, seq, _HDR.size)]
    mark("verdict_exchange")
    rc = max(max(f[0] for f in fin), fin[0][4])
    if rc:
        return rc, False
    return 0, bool(fin[0][5])
