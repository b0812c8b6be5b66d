#!/usr/bin/env python3
"""Checks the inversion-free final exponentiation used by csrc/mpair.cuh against the Python model.

With u = |x|, m = f^(p^2+1) and H+ - H- = 3 (p^4 - p^2 + 1) / r,
    H+ = (u+1)^2 (p u^2 + p^3 + u) + 3,    H- = (u+1)^2 (p + u^3 + u p^2),
the value f^(3 (p^12-1)/r) equals conj(X+) X- / (X+ conj(X-)) for X+- = m^(H+-); the device compares
conj(X+) X- with X+ conj(X-) instead of computing an inverse.  Run: python tools/check_final_exp.py
"""
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.pymodel import bls12_381 as m  # noqa: E402

P, R, U = m.P, m.R, m.X_ABS


def split_check(f):
    """(conj(X+) X-, X+ conj(X-)) computed exactly as mp_final_check does."""
    mm = m.f12_mul(m.f12_frob(f, 2), f)
    m1 = m.f12_pow(mm, U)
    m2 = m.f12_pow(m1, U)
    n = m.f12_mul(m.f12_mul(m2, m.f12_sqr(m1)), mm)
    n1 = m.f12_pow(n, U)
    n2 = m.f12_pow(n1, U)
    n3 = m.f12_pow(n2, U)
    m3 = m.f12_mul(m.f12_sqr(mm), mm)
    xp = m.f12_mul(m.f12_mul(m.f12_frob(n2, 1), m.f12_frob(n, 3)), m.f12_mul(n1, m3))
    xm = m.f12_mul(m.f12_mul(m.f12_frob(n, 1), n3), m.f12_frob(n1, 2))
    return m.f12_mul(m.f12_conj(xp), xm), m.f12_mul(xp, m.f12_conj(xm))


def main():
    a = (U + 1) ** 2
    hp = a * (P * U * U + P**3 + U) + 3
    hm = a * (P + U**3 + U * P * P)
    assert hp - hm == 3 * (P**4 - P**2 + 1) // R
    rnd = random.Random(5)
    f = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(6)]
    lhs, rhs = split_check(f)
    assert m.f12_mul(rhs, m.f12_pow(m.final_exp(f), 3)) == lhs
    print("final exponentiation identity ok")


if __name__ == "__main__":
    main()
