#!/usr/bin/env python3
"""Generates tests/golden/vectors_small.json with the pure-Python model (oracle/pymodel): a small batch,
every stage artefact, negative encodings with their status bytes, and Fp12 known answers.  The upstream
reference has no golden vectors (LICENSE only), so these model-generated vectors are what pins the C++
oracle and, through the same files, the CUDA library.  Run from the repo root; deterministic."""
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pymodel import bls12_381 as b  # noqa: E402
from oracle.pymodel import kzg_model as k  # noqa: E402
from tests.helpers import f12_bytes, negative_g1_encodings, rand_g1  # noqa: E402


def main():
    out = {}
    seed, n = 0x4B5A47AA, 5
    C, Z, Y, PI = k.gen_instance_shortcut(seed, n)
    art = k.batch_artifacts(C, Z, Y, PI, n)
    g2t = b.g2_mul(k.TAU, b.G2)
    out["batch"] = {
        "seed": seed, "n": n, "C": C.hex(), "z": Z.hex(), "y": Y.hex(), "pi": PI.hex(),
        "root": art["root"].hex(), "r": [f"{v:032x}" for v in art["r"]],
        **{key: b.g1_affine_bytes(art[key]).hex() for key in ("S1", "S2", "S3", "A", "B")},
        "sum_ry": f"{art['sum_ry']:064x}", "verdict": k.verdict_pairing(art, g2t),
        "verdict_tau_shortcut": k.verdict_tau_shortcut(art),
    }
    j = k.plant_index(seed, n)
    PIbad = k.plant_invalid(PI, j)
    art2 = k.batch_artifacts(C, Z, Y, PIbad, n)
    out["batch_planted"] = {"index": j, "pi": PIbad.hex(), "A": b.g1_affine_bytes(art2["A"]).hex(),
                            "B": b.g1_affine_bytes(art2["B"]).hex(), "verdict": k.verdict_pairing(art2, g2t)}
    a1 = k.batch_artifacts(C[:48], Z[:32], Y[:32], PI[:48], 1, single=True)
    out["single"] = {"A": b.g1_affine_bytes(a1["A"]).hex(), "B": b.g1_affine_bytes(a1["B"]).hex(),
                     "verdict": k.verdict_pairing(a1, g2t)}
    rnd = random.Random(2026)
    out["encodings"] = [{"in": enc.hex(), "status": st, "affine": b.g1_affine_bytes(b.g1_decompress(enc)[1]).hex()}
                        for enc, st in negative_g1_encodings(rnd)]
    A, B = rand_g1(rnd), rand_g1(rnd)
    gt = b.final_exp(b.f12_mul(b.miller_loop(A, b.G2), b.miller_loop(B, g2t)))
    out["pairing"] = {"A": b.g1_affine_bytes(A).hex(), "B": b.g1_affine_bytes(B).hex(),
                      "gt_cubed": f12_bytes(b.f12_mul(b.f12_mul(gt, gt), gt)).hex(),
                      "note": "f^(3(p^12-1)/r) of miller(A,G2)*miller(B,[tau]G2); Fp12 as 6 Fp2 coefficients of w^0..w^5 (c0|c1)"}
    pts = [rand_g1(rnd) for _ in range(12)] + [None]
    ks = [rnd.randrange(b.R) for _ in range(13)]
    acc = None
    for p_, k_ in zip(pts, ks):
        acc = b.g1_add(acc, b.g1_mul(k_, p_))
    out["msm"] = {"points": b"".join(b.g1_affine_bytes(p_) for p_ in pts).hex(), "scalars": b"".join(x.to_bytes(32, "big") for x in ks).hex(),
                  "result": b.g1_affine_bytes(acc).hex()}
    path = ROOT / "tests" / "golden" / "vectors_small.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
