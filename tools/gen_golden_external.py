#!/usr/bin/env python3
"""Generates tests/golden/external_setup_openings.json: openings of degree-1 polynomials under a setup whose
tau nobody in this repo knows.

PROVENANCE -- read before citing.  The three hex strings below were typed in from the author's memory; they were
NOT read from /root/reference (which holds a LICENSE only), from a file in this image, or from the network (there
is none).  The author believes T1/T2 are [tau]G1 / [tau]G2 of a public KZG ceremony, but that attribution is
unverified and must not be quoted as fact.  What IS verified, by this script every time it runs:
  * each string decompresses to a point of the prime-order subgroup (a random string passes with p ~ 2^-127),
  * e(T1, G2) == e(G1, T2), i.e. T1 and T2 hide the same unknown scalar.
That makes the pair a usable structured reference string with an unknown tau, which is what the test needs: the
tau-shortcut cannot be used, so acceptance has to come from the real pairing.  It does not pin the oracle to the
upstream reference (DESIGN.md keeps "parity unpinned").

p_i(X) = a_i + b_i X  =>  C_i = a_i G1 + b_i T1,  y_i = a_i + b_i z_i,  pi_i = b_i G1  (quotient is the constant b_i).
Run from the repo root; deterministic."""
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.pymodel import bls12_381 as b  # noqa: E402

T1_HEX = "ad3eb50121139aa34db1d545093ac9374ab7bca2c0f3bf28e27c8dcd8fc7cb42d25926fc0c97b336e9f0fb35e5a04c81"
T2_HEX = ("b5bfd7dd8cdeb128843bc287230af38926187075cbfbefa81009a2ce615ac53d2914e5870cb452d2afaaab24f3499f72"
          "185cbfee53492714734429b7b38608e23926c911cceceac9a36851477ba4c60b087041de621000edc98edada20c1def2")
# further recalled strings of unknown origin; only claim: they decompress to G1 subgroup points
EXTRA_G1_HEX = [
    "854262641262cb9e056a8512808ea6864d903dbcad713fd6da8dddfa5ce40d85612c912063ace060ed8c4bf005bab839",
    "a0413c0dcafec6dbc9f47d66785cf1e8c981044f7d13cfe3e4fcbb71b5408dfde6312493cb3c1d30516cb3ca88c03654",
    "8b997fb25730d661918371bb41f2a6e899cac23f04fc5365800b75433c0a953250e15e7a98fb5ca5cc56a8cd34c20c57",
]


def main():
    st, T1 = b.g1_decompress(bytes.fromhex(T1_HEX), subgroup="slow")[:2]
    assert st == b.ST_OK
    T2 = b.g2_decompress(bytes.fromhex(T2_HEX))
    assert b.g2_mul(b.R, T2) is None
    assert b.pairing_product_is_one([(T1, b.G2), (b.g1_neg(b.G1), T2)]), "T1/T2 do not share a scalar"
    extra = []
    for h in EXTRA_G1_HEX:
        st, pt = b.g1_decompress(bytes.fromhex(h), subgroup="slow")[:2]
        assert st == b.ST_OK
        extra.append({"in": h, "affine": b.g1_affine_bytes(pt).hex()})
    rnd = random.Random(0xE87)
    n = 24
    C, Z, Y, PI = b"", b"", b"", b""
    for i in range(n):
        a_, b_, z_ = rnd.randrange(b.R), rnd.randrange(1, b.R), rnd.randrange(b.R)
        if i == 3:
            z_ = 0
        if i == 4:
            a_ = 0
        C += b.g1_compress(b.g1_add(b.g1_mul(a_, b.G1), b.g1_mul(b_, T1)))
        Z += z_.to_bytes(32, "big")
        Y += ((a_ + b_ * z_) % b.R).to_bytes(32, "big")
        PI += b.g1_compress(b.g1_mul(b_, b.G1))
    out = {
        "provenance": "T1/T2/extra recalled from memory by the author, origin unverified; accepted only because the "
                      "subgroup checks and e(T1,G2)==e(G1,T2) hold (tools/gen_golden_external.py re-checks both)",
        "g1_monomial": (b.g1_compress(b.G1) + bytes.fromhex(T1_HEX)).hex(),
        "g2_monomial": (b.g2_compress(b.G2) + bytes.fromhex(T2_HEX)).hex(),
        "T1_affine": b.g1_affine_bytes(T1).hex(),
        "extra_g1": extra,
        "n": n, "C": C.hex(), "z": Z.hex(), "y": Y.hex(), "pi": PI.hex(),
    }
    path = ROOT / "tests" / "golden" / "external_setup_openings.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
