"""One process per GPU (torchrun) over the shard-level C ABI of the CUDA library, on every physical GPU of the box:
the verdict and the pairing inputs A, B, sum r_i y_i, root equal the single-device batch and the oracle; the oracle's
combine accepts the CUDA shards' pairing terms (cross-library); planted wrong proof / off-subgroup point on the last
rank are rejected on every rank.  Skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KZGB_ROOT"])
import torch
import torch.distributed as dist
from kzg_batch_verification_scheme_b200.api import KzgLib, load, TERMS_BYTES
from kzg_batch_verification_scheme_b200.sharded import HostMailbox, sharded_verify
from oracle.pymodel import bls12_381 as b
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="gloo", rank=rank, world_size=world)
lib = load()
n_local, seed = 16384 + 128 * 3, 0x4B5A4741
n_total = n_local * world
ctx = lib.test_context(devices=[local], n_max=n_total)
box = HostMailbox(dist, rank, world, n_local)
C, Z, Y, PI = ctx.synth_instance(seed, rank * n_local, n_local)
ref = None
if rank == 0:
    full = ctx.synth_instance(seed, 0, n_total)
    assert ctx.verify_kzg_proof_batch(*full, n_total) == (0, True)
    ref = ctx.last_artifacts()
    olib = KzgLib(os.path.join(os.environ["KZGB_ROOT"], "oracle", "libkzgb_oracle.so"))
    octx = olib.test_context()
    assert octx.verify_kzg_proof_batch(*full, n_total) == (0, True)
    oref = octx.last_artifacts()
    for key in ("A", "B", "sum_ry", "root"):
        assert ref[key] == oref[key], key
for mode in ("terms", "partials"):
    for _ in range(2):
        assert sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_local, box=box, mode=mode) == (0, True), mode
        if rank == 0:
            art = ctx.last_artifacts()
            for key in ("A", "B", "sum_ry"):
                assert art[key] == ref[key], (mode, key)
# device-resident inputs
bufs = [torch.frombuffer(bytearray(x), dtype=torch.uint8).cuda() for x in (C, Z, Y, PI)]
torch.cuda.synchronize()
assert sharded_verify(ctx, dist, rank, world, *[t.data_ptr() for t in bufs], n_local, on_device=True,
                      stream=torch.cuda.current_stream().cuda_stream, box=box) == (0, True)
# cross-library: the oracle's combine on the CUDA shards' terms
digs = [None] * world
rc, dig, _ = ctx.shard_phase1(0, C, Z, Y, PI, n_local)
dist.all_gather_object(digs, dig)
root = ctx.fs_root(b"".join(digs), n_total)
rc, terms = ctx.shard_phase2_terms(0, root, rank * n_local)
assert rc == 0 and len(terms) == TERMS_BYTES and ctx.shard_finish(0) == (0, 0, 0)
allt = [None] * world
dist.all_gather_object(allt, terms)
if rank == 0:
    assert root == ref["root"]
    assert octx.combine_verify_terms(b"".join(allt)) == (0, True)
    oa = octx.last_artifacts()
    assert oa["A"] == ref["A"] and oa["B"] == ref["B"] and oa["sum_ry"] == ref["sum_ry"]
    assert ctx.combine_verify_terms(b"".join(allt)) == (0, True)
# planted wrong proof on the last rank: rejected everywhere
PIb = PI[:48] + PI[:48] + PI[96:] if rank == world - 1 else PI
assert sharded_verify(ctx, dist, rank, world, C, Z, Y, PIb, n_local, box=box) == (0, False)
# off-subgroup commitment on the last rank: BADARGS everywhere
Cb = C[:48 * 7] + b.g1_compress((0, 2)) + C[48 * 8:] if rank == world - 1 else C
assert sharded_verify(ctx, dist, rank, world, Cb, Z, Y, PI, n_local, box=box) == (1, False)
assert sharded_verify(ctx, dist, rank, world, C, Z, Y, PI, n_local, box=box) == (0, True)
box.close()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gpu_one_process_per_gpu_terms_exchange(gpu_lib, oracle_lib, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 physical GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KZGB_ROOT=str(ROOT), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == world
