"""Robustness of the CUDA library beyond single calls: context life cycle (no device-memory leak), contexts used from
concurrent host threads, batches of exactly n_max proofs, the deepest pipeline."""
import threading

import pytest

pytestmark = pytest.mark.gpu


def test_gpu_context_life_cycle_does_not_leak(gpu_lib):
    import torch
    ctx = gpu_lib.test_context(n_max=1 << 14)
    ctx.pipeline_init(2)
    ctx.close()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(12):
        ctx = gpu_lib.test_context(n_max=1 << 14, cells=True)
        assert ctx.pipeline_init(3) == 0
        C, Z, Y, PI = ctx.synth_instance(0x4B5A47D0, 0, 200)
        assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, 200) == (0, True)
        rc, t = ctx.verify_kzg_proof_batch_submit(C, Z, Y, PI, 200)
        assert rc == 0 and ctx.verify_kzg_proof_batch_wait(t) == (0, True)
        ctx.close()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, (free0, free1)          # nothing but allocator noise


def test_gpu_two_contexts_from_two_threads(gpu_lib, oracle_ctx):
    n = 3000
    inst = [oracle_ctx.synth_instance(0x4B5A47E0 + k, 0, n) for k in range(2)]
    bad = [(c, z, y, p[48:96] + p[:48] + p[96:]) for (c, z, y, p) in inst]
    ctxs = [gpu_lib.test_context(n_max=4096) for _ in range(2)]
    errors = []

    def work(k):
        try:
            for it in range(25):
                assert ctxs[k].verify_kzg_proof_batch(*inst[k], n) == (0, True)
                assert ctxs[k].verify_kzg_proof_batch(*bad[k], n) == (0, False)
        except Exception as ex:            # noqa: BLE001
            errors.append(repr(ex))
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("n_max", [1, 2, 127, 128, 129, 5000])
def test_gpu_batch_of_exactly_n_max(gpu_lib, oracle_ctx, n_max):
    ctx = gpu_lib.test_context(n_max=n_max)
    C, Z, Y, PI = oracle_ctx.synth_instance(0x4B5A47F0 + n_max, 0, n_max)
    assert ctx.verify_kzg_proof_batch(C, Z, Y, PI, n_max) == oracle_ctx.verify_kzg_proof_batch(C, Z, Y, PI, n_max) == (0, True)
    a1, a2 = ctx.last_artifacts(), oracle_ctx.last_artifacts()
    for key in ("S1", "S2", "S3", "A", "B", "sum_ry", "root"):
        assert a1[key] == a2[key], key
    big = oracle_ctx.synth_instance(1, 0, n_max + 1)
    assert ctx.verify_kzg_proof_batch(*big, n_max + 1) == (1, False)         # larger than the workspaces: refused, not a crash
    ctx.close()


def test_gpu_deepest_pipeline(gpu_lib, oracle_ctx):
    from tests import parity_suite as ps
    ctx = gpu_lib.test_context(n_max=2048)
    ps.check_pipeline(ctx, oracle_ctx, depth=8, sizes=(100, 2048, 3, 700, 64, 1, 129, 1500, 5, 2047, 300, 40), seed=0x4B5A4801)
    ctx.close()
