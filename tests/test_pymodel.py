"""Pure-Python model self-checks: known answers of SURVEY.md Appendix A + algebraic pairing tests."""
import random

import pytest

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k


def test_constants_and_encodings():
    assert b.g1_on_curve(b.G1) and b.g2_on_curve(b.G2)
    assert b.g1_mul(b.R, b.G1) is None and b.g2_mul(b.R, b.G2) is None
    assert b.H1 == 0x396C8C005555E1568C00AAAB0000AAAB
    assert b.g1_compress(b.G1).hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                                          "6c55e83ff97a1aeffb3af00adb22c6bb")
    assert b.g1_compress(b.g1_mul(2, b.G1)).hex() == ("a572cbea904d67468808c8eb50a9450c9721db309128012543902d0a"
                                                       "c358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e")
    assert b.g1_compress(b.g1_neg(b.G1)).hex()[:4] == "b7f1"
    assert b.g1_compress(None) == bytes([0xC0]) + bytes(47)
    assert b.g2_compress(b.G2).hex().startswith("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049")
    assert b.g2_decompress(b.g2_compress(b.G2)) == b.G2
    assert b.BETA == 0x5F19672FDF76CE51BA69C6076A0F77EADDB3A93BE6F89688DE17D813620A00022E01FFFFFFFEFFFE
    assert pow(b.BETA, 3, b.P) == 1


def test_subgroup_fast_equals_slow():
    rnd = random.Random(7)
    assert b.g1_in_subgroup_fast(b.G1)
    assert b.g1_on_curve((0, 2)) and not b.g1_in_subgroup_fast((0, 2)) and not b.g1_in_subgroup_slow((0, 2))
    n = 0
    while n < 6:
        x = rnd.randrange(b.P)
        y = b.fp_sqrt((x ** 3 + 4) % b.P)
        if y is None:
            continue
        n += 1
        pt = (x, y)
        assert b.g1_in_subgroup_fast(pt) == b.g1_in_subgroup_slow(pt)
        cleared = b.g1_mul(b.H1, pt)
        assert b.g1_in_subgroup_fast(cleared) and b.g1_in_subgroup_slow(cleared)
        tors = b.g1_mul(b.R, pt)           # lies in the cofactor part
        if tors is not None:
            mixed = b.g1_add(b.G1, tors)
            assert not b.g1_in_subgroup_fast(mixed) and not b.g1_in_subgroup_slow(mixed)


def test_pairing_algebra_and_final_exp_identity():
    e1 = b.pairing(b.G1, b.G2)
    assert e1 != b.f12_one() and b.f12_pow(e1, b.R) == b.f12_one()
    assert b.pairing(b.g1_mul(5, b.G1), b.g2_mul(7, b.G2)) == b.f12_pow(e1, 35)
    assert b.f12_mul(b.pairing(b.G1, b.g2_neg(b.G2)), e1) == b.f12_one()
    x, p = b.X, b.P
    assert 3 * ((p ** 4 - p ** 2 + 1) // b.R) == (x - 1) ** 2 * (x + p) * (x ** 2 + p ** 2 - 1) + 3
    rnd = random.Random(1)
    a = [(rnd.randrange(p), rnd.randrange(p)) for _ in range(6)]
    assert b.f12_frob(a, 1) == b.f12_pow(a, p)
    assert b.f12_mul(a, b.f12_inv(a)) == b.f12_one()


def test_kzg_semantics_small():
    C, Z, Y, PI = k.gen_instance_shortcut(0x4B5A4701, 4)
    art = k.batch_artifacts(C, Z, Y, PI, 4)
    g2t = b.g2_mul(k.TAU, b.G2)
    assert art["ret"] == 0 and k.verdict_tau_shortcut(art) and k.verdict_pairing(art, g2t)
    j = k.plant_index(0x4B5A4701, 4)
    art2 = k.batch_artifacts(C, Z, Y, k.plant_invalid(PI, j), 4)
    assert art2["ret"] == 0 and not k.verdict_tau_shortcut(art2) and not k.verdict_pairing(art2, g2t)
    C, Z, Y, PI = k.gen_instance_poly(5, 2, 16)
    assert k.verdict_tau_shortcut(k.batch_artifacts(C, Z, Y, PI, 2))
    art1 = k.batch_artifacts(C[:48], Z[:32], Y[:32], PI[:48], 1, single=True)
    assert k.verdict_tau_shortcut(art1)


def _signed_digits(r, c, nbits=128):
    """Digit recoding of csrc/msm.cuh (msm_digits_body): signed windows, unsigned top window absorbing the carry."""
    W = (nbits + c - 1) // c
    out, carry = [], 0
    for w in range(W):
        width = c if w < W - 1 else nbits - c * (W - 1)
        v = ((r >> (c * w)) & ((1 << width) - 1)) + carry
        if w < W - 1 and v > (1 << (c - 1)):
            out.append(v - (1 << c)); carry = 1
        else:
            out.append(v); carry = 0
    assert sum(d << (c * w) for w, d in enumerate(out)) == r
    return out


@pytest.mark.parametrize("c", [3, 8, 13, 16])
def test_slice_coefficients_of_the_batched_subgroup_check_are_fair_coins(c):
    """DESIGN.md "Batched subgroup check": a point enters slice (w, b) with coefficient sign * bit_b(|digit|) and slice
    (w, all) with its sign.  Soundness needs every coefficient value to have probability <= 1/2 (exact: of the 2^c
    equally likely window values, half have bit b of the magnitude clear, half are positive) and the 128 coefficients
    of one point to be independent.  Empirical check over random 128-bit challenges."""
    import random
    rnd = random.Random(1000 + c)
    nbits, N = 128, 6000
    W = (nbits + c - 1) // c
    tb = nbits - c * (W - 1)
    slices = [(w, b) for w in range(W - 1) for b in range(c)] + [(W - 1, b) for b in range(tb)]
    assert len(slices) == nbits
    counts = [dict() for _ in slices]
    zero_pairs = 0
    s0, s1 = 0, len(slices) // 2                     # two slices of different windows
    z0 = z1 = 0
    for _ in range(N):
        d = _signed_digits(rnd.getrandbits(nbits), c)
        coef = []
        for w, b in slices:
            m, sgn = abs(d[w]), (1 if d[w] > 0 else -1 if d[w] < 0 else 0)
            if w < W - 1 and b == c - 1:
                v = sgn                                # the slice "all"
            else:
                v = sgn * ((m >> b) & 1)
            coef.append(v)
        for k, v in enumerate(coef):
            counts[k][v] = counts[k].get(v, 0) + 1
        a, bq = coef[s0] == 0, coef[s1] == 0
        z0 += a; z1 += bq; zero_pairs += a and bq
    slack = 4 * (0.25 / N) ** 0.5                       # four standard deviations
    for k, cnt in enumerate(counts):
        assert max(cnt.values()) / N <= 0.5 + slack, (slices[k], cnt)
    assert abs(zero_pairs / N - (z0 / N) * (z1 / N)) < 0.03          # no visible dependence between windows
