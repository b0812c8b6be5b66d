#!/usr/bin/env python3
"""Soak run on a B200: the randomised differential test of tests/parity_suite.py at larger and odd batch sizes (covering the
head-mode host path, the early-S1 range, window widths 9..13) against the oracle, plain and EIP-4844 transcript.
Usage: python tools/gpu_soak.py [trials]"""
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from kzg_batch_verification_scheme_b200.api import KzgLib, load  # noqa: E402
from tests import parity_suite as ps  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
gpu = load().test_context(n_max=1 << 17)
orc = KzgLib(ROOT / "oracle" / "libkzgb_oracle.so").test_context()
t0 = time.time()
sizes = (4095, 4096, 9000, 16383, 16384, 16385, 24576, 40000, 57343, 57344, 65535, 65536, 65537, 70001, 90112, 90113, 100000)
ps.check_random_differential(gpu, orc, sizes=sizes, trials=trials, pool=100000)
print(f"plain transcript: {trials} randomised batches up to 100000 proofs equal the oracle ({time.time() - t0:.0f} s)", flush=True)
# EIP-4844 mode: same corruptions, smaller sizes (per-point subgroup checks, 255-bit sums)
rnd = random.Random(77)
C0, Z0, Y0, PI0 = orc.synth_instance(0x4B5A4740, 0, 20000)
t0 = time.time()
for trial in range(trials // 2):
    n = rnd.choice([1, 2, 63, 1000, 4096, 8191, 20000])
    off = rnd.randrange(0, 20000 - n + 1)
    arrs = [bytearray(C0[48 * off:48 * (off + n)]), bytearray(Z0[32 * off:32 * (off + n)]), bytearray(Y0[32 * off:32 * (off + n)]),
            bytearray(PI0[48 * off:48 * (off + n)])]
    kind = rnd.choice(["none", "flip", "swap", "highbit"])
    if kind == "flip":
        a = rnd.randrange(4)
        arrs[a][rnd.randrange(len(arrs[a]))] ^= 1 << rnd.randrange(8)
    elif kind == "swap" and n >= 2:
        i, j = rnd.sample(range(n), 2)
        arrs[3][48 * i:48 * i + 48], arrs[3][48 * j:48 * j + 48] = arrs[3][48 * j:48 * j + 48], arrs[3][48 * i:48 * i + 48]
    elif kind == "highbit":
        arrs[rnd.choice([1, 2])][32 * rnd.randrange(n)] |= 0x80
    args = [bytes(x) for x in arrs] + [n]
    r1, r2 = gpu.verify_kzg_proof_batch_eip4844(*args), orc.verify_kzg_proof_batch_eip4844(*args)
    assert r1 == r2, (trial, kind, n, r1, r2)
    if r1[0] == 0:
        a1, a2 = gpu.last_artifacts(), orc.last_artifacts()
        for key in ("A", "B", "sum_ry", "root"):
            assert a1[key] == a2[key], (trial, kind, key)
print(f"EIP-4844 transcript: {trials // 2} randomised batches up to 20000 proofs equal the oracle ({time.time() - t0:.0f} s)", flush=True)
