"""World-size-2 (and 3) test of the one-process-per-GPU host logic on CPU: gloo backend, the CPU oracle as the
library behind the shard-level C ABI.  Verdict and artefacts must equal the single-process batch."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KZGB_ROOT"])
import torch.distributed as dist
from kzg_batch_verification_scheme_b200.api import KzgLib
from kzg_batch_verification_scheme_b200.sharded import sharded_verify
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group(backend="gloo", rank=rank, world_size=world)
lib = KzgLib(os.path.join(os.environ["KZGB_ROOT"], "oracle", "libkzgb_oracle.so"))
ctx = lib.test_context()
ctx.set_threads(2)
from kzg_batch_verification_scheme_b200.sharded import HostMailbox
n_local, seed = 1024, 0x4B5A4705
mailbox = HostMailbox(dist, rank, world, 2 * n_local, tag="test%d" % world)
full = lib.test_context()
for _ in range(3):
    mailbox.barrier()
for box, mode in ((None, "terms"), (None, "partials"), (mailbox, "terms"), (mailbox, "partials")):
    C, Z, Y, PI = ctx.synth_instance(seed, rank * n_local, n_local)
    rc, ok = sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_local, box=box, mode=mode)
    assert (rc, ok) == (0, True), (rc, ok, mode)              # every rank returns the verdict
    if rank == 0:
        art = ctx.last_artifacts()
        Cf, Zf, Yf, PIf = full.synth_instance(seed, 0, n_local * world)
        assert full.verify_kzg_proof_batch(Cf, Zf, Yf, PIf, n_local * world) == (0, True)
        ref = full.last_artifacts()
        assert art["A"] == ref["A"] and art["B"] == ref["B"] and art["sum_ry"] == ref["sum_ry"]
    # a wrong proof on the LAST rank must flip the verdict everywhere
    if rank == world - 1:
        PI = PI[:48] + PI[:48] + PI[96:]
    assert sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_local, box=box, mode=mode) == (0, False)
    # a malformed element on rank 1 is BADARGS everywhere
    if rank == min(1, world - 1):
        Z = bytes([0xFF]) * 32 + Z[32:]
    rc, ok = sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_local, box=box, mode=mode)
    assert (rc, ok) == (1, False), rc
# shards of different sizes (the last one ragged): offsets are the prefix sums of the gathered sizes
sizes = [256 * (r + 1) for r in range(world - 1)] + [333]
off, n_r, n_total = sum(sizes[:rank]), sizes[rank], sum(sizes)
C, Z, Y, PI = ctx.synth_instance(seed + 1, off, n_r)
assert sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_r, box=mailbox) == (0, True)
if rank == 0:
    art = ctx.last_artifacts()
    Cf, Zf, Yf, PIf = full.synth_instance(seed + 1, 0, n_total)
    assert full.verify_kzg_proof_batch(Cf, Zf, Yf, PIf, n_total) == (0, True)
    ref = full.last_artifacts()
    assert art["A"] == ref["A"] and art["B"] == ref["B"] and art["sum_ry"] == ref["sum_ry"]
assert sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_r) == (0, True)          # same over gloo tensors
# a shard boundary off the 128-proof chunk grid is refused on every rank
if world > 1:
    n_bad = 100 if rank == 0 else 128
    Cb, Zb, Yb, PIb = ctx.synth_instance(seed + 2, 0, n_bad)
    assert sharded_verify(ctx, dist, rank, world, Cb, Zb, Yb, PIb, n_bad, box=mailbox, n_max_local=2 * n_local) == (1, False)
mailbox.close()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_verify_gloo(oracle_lib, world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KZGB_ROOT=str(ROOT), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
