"""Shared test helpers: deterministic operand makers and negative-input constructors (pure Python)."""
import hashlib
import random

from oracle.pymodel import bls12_381 as b
from oracle.pymodel import kzg_model as k

P, R = b.P, b.R


def fp_bytes(v):
    return (v % P).to_bytes(48, "big")


def fr_bytes(v):
    return (v % R).to_bytes(32, "big")


def rand_fp(rnd):
    return rnd.randrange(P)


def edge_fps():
    return [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 2 ** 380, 2 ** 381 - 1 - (2 ** 381 - 1 >= P) * (2 ** 381 - P),
            (1 << 384) % P, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, P - 0xFFFFFFFF]


def rand_g1(rnd):
    return b.g1_mul(rnd.randrange(1, R), b.G1)


def rand_curve_point(rnd):
    """Random point of E(Fp), almost surely outside G1."""
    while True:
        x = rnd.randrange(P)
        y = b.fp_sqrt((x ** 3 + 4) % P)
        if y is not None:
            return (x, y if rnd.random() < 0.5 else P - y)


def f12_bytes(a):
    return b"".join(fp_bytes(c[0]) + fp_bytes(c[1]) for c in a)


def f12_from_bytes(bs):
    return [(int.from_bytes(bs[96 * i:96 * i + 48], "big"), int.from_bytes(bs[96 * i + 48:96 * i + 96], "big"))
            for i in range(6)]


def negative_g1_encodings(rnd):
    """(bytes, expected status) covering every class of App. B.2."""
    out = []
    good = b.g1_compress(rand_g1(rnd))
    out.append((good, 0))
    out.append((b.g1_compress(None), 0))
    out.append((bytes([good[0] & 0x7F]) + good[1:], 1))                 # compressed bit clear
    out.append((bytes([0xC0]) + bytes(46) + b"\x01", 1))                # infinity with junk
    out.append((bytes([0xE0]) + bytes(47), 1))                          # infinity with sign bit
    out.append((bytes([0x40]) + bytes(47), 1))                          # infinity without compressed bit
    xb = bytearray(P.to_bytes(48, "big")); xb[0] |= 0x80
    out.append((bytes(xb), 2))                                           # x == p
    xb = bytearray((P + 5).to_bytes(48, "big")); xb[0] |= 0x80
    out.append((bytes(xb), 2))                                           # x > p
    xb = bytearray((2 ** 381 - 1).to_bytes(48, "big")); xb[0] |= 0x80
    out.append((bytes(xb), 2))
    while True:                                                          # x^3+4 non-residue
        x = rnd.randrange(P)
        if b.fp_sqrt((x ** 3 + 4) % P) is None:
            xb = bytearray(x.to_bytes(48, "big")); xb[0] |= 0x80
            out.append((bytes(xb), 3))
            break
    out.append((b.g1_compress((0, 2)), 4))                               # order-3 point
    out.append((b.g1_compress((0, P - 2)), 4))
    out.append((b.g1_compress(rand_curve_point(rnd)), 4))                # random curve point
    tors = b.g1_mul(R, rand_curve_point(rnd))                            # cofactor-torsion point
    out.append((b.g1_compress(tors), 4 if tors is not None else 0))
    out.append((b.g1_compress(b.g1_add(rand_g1(rnd), tors)), 4))         # G1 + torsion
    for q in (11, 10177, 859267, 52437899):                              # order-q points
        pt = b.g1_mul(b.H1 * R // (q * q), rand_curve_point(rnd))
        while pt is None:
            pt = b.g1_mul(b.H1 * R // (q * q), rand_curve_point(rnd))
        out.append((b.g1_compress(pt), 4))
    return out
