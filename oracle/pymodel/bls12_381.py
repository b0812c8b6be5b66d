"""Pure-Python model of BLS12-381 (stdlib big ints only).  TEST INFRASTRUCTURE.

Independent ground truth for the C++ oracle (oracle/*.hpp) and, through it, for the CUDA path.
Everything is derived from the curve parameter x = -0xd201000000010000 and the two generator
coordinates; nothing is transcribed from a library.

Reference citation: the upstream reference (/root/reference) holds only LICENSE:1-201 -- there is
no implementation to follow.  Semantics follow BASELINE.json:5 (north_star) and SURVEY.md App. B.

Deliberately naive: affine arithmetic, double-and-add, a textbook ate pairing evaluated in
E(Fp12) through the untwist map, and a final exponentiation by plain pow((p^12-1)/r).
"""
from __future__ import annotations

X_ABS = 0xD201000000010000          # |x|; the BLS parameter is x = -X_ABS
X = -X_ABS
R = X**4 - X**2 + 1                 # group order
P = (X - 1) ** 2 * R // 3 + X       # base field
H1 = (X - 1) ** 2 // 3              # G1 cofactor
B = 4

assert P == 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
assert R == 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
assert P % 4 == 3 and P % 6 == 1

G1X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2X = (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
       0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E)
G2Y = (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
       0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)

BETA = pow(2, (P - 1) // 3, P)      # primitive cube root of unity in Fp


# ----------------------------------------------------------------------------- Fp
def fp_inv(a):
    return pow(a, P - 2, P)


def fp_sqrt(a):
    """Square root for p = 3 mod 4, or None."""
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


# ----------------------------------------------------------------------------- Fp2 = Fp[u]/(u^2+1)
def f2(a, b=0):
    return (a % P, b % P)


F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (1, 1)                          # v^3 = xi = 1+u


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_muls(a, s):
    return (a[0] * s % P, a[1] * s % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    n = fp_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * n % P, -a[1] * n % P)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def f2_sqrt(a):
    """Any square root in Fp2 (p = 3 mod 4), or None.  Brute-force-simple: via norm."""
    if a == F2_ZERO:
        return F2_ZERO
    a1 = f2_pow(a, (P - 3) // 4)
    alpha = f2_mul(f2_sqr(a1), a)
    x0 = f2_mul(a1, a)
    if alpha == (P - 1, 0):
        x = f2_mul((0, 1), x0)
    else:
        b = f2_pow(f2_add(F2_ONE, alpha), (P - 1) // 2)
        x = f2_mul(b, x0)
    return x if f2_sqr(x) == a else None


# ----------------------------------------------------------------------------- Fp12 as Fp2[w]/(w^6 - xi)
# element = list of 6 Fp2 coefficients of w^0..w^5.  (Tower view: v = w^2, Fp6 = {w^0,w^2,w^4},
# Fp12 = Fp6 + Fp6*w.)
def f12_one():
    return [F2_ONE] + [F2_ZERO] * 5


def f12_from_f2(a):
    return [a] + [F2_ZERO] * 5


def f12_add(a, b):
    return [f2_add(x, y) for x, y in zip(a, b)]


def f12_sub(a, b):
    return [f2_sub(x, y) for x, y in zip(a, b)]


def f12_mul(a, b):
    t = [F2_ZERO] * 11
    for i in range(6):
        if a[i] == F2_ZERO:
            continue
        for j in range(6):
            t[i + j] = f2_add(t[i + j], f2_mul(a[i], b[j]))
    return [f2_add(t[k], f2_mul(XI, t[k + 6])) if k < 5 else t[k] for k in range(6)]


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    """a^(p^6): w -> -w."""
    return [a[k] if k % 2 == 0 else f2_neg(a[k]) for k in range(6)]


_FROB_GAMMA = [[f2_pow(XI, k * (P**j - 1) // 6) for k in range(6)] for j in range(0, 4)]


def f12_frob(a, j=1):
    """a^(p^j), j in 1..3."""
    out = []
    for k in range(6):
        c = a[k] if j % 2 == 0 else f2_conj(a[k])
        out.append(f2_mul(c, _FROB_GAMMA[j][k]))
    return out


def f12_pow(a, e):
    r = f12_one()
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


def f12_inv(a):
    """Inverse via a * conj-products: a^-1 = a^(p^12-2) is too slow; use the norm to Fp6/Fp2 chain
    expressed through Frobenius: N = a * a^(p^6) lies in Fp6; then N6 = N*N^(p^2)*N^(p^4) in Fp2."""
    abar = f12_conj(a)
    n6 = f12_mul(a, abar)                       # in Fp6 (only even w-powers)
    n6p2 = f12_frob(n6, 2)
    n6p4 = f12_frob(n6p2, 2)
    t = f12_mul(n6p2, n6p4)
    n2 = f12_mul(n6, t)                         # in Fp2
    assert all(c == F2_ZERO for c in n2[1:])
    s = f2_inv(n2[0])
    return [f2_mul(c, s) for c in f12_mul(abar, t)]


# ----------------------------------------------------------------------------- curves (affine, None = infinity)
class _Field:
    pass


def _mk_curve(add, sub, mul, inv, neg, zero, three_mul, b):
    def ec_add(p1, p2):
        if p1 is None:
            return p2
        if p2 is None:
            return p1
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            if y1 == y2 and y1 != zero:
                lam = mul(three_mul(mul(x1, x1)), inv(add(y1, y1)))
            else:
                return None
        else:
            lam = mul(sub(y2, y1), inv(sub(x2, x1)))
        x3 = sub(sub(mul(lam, lam), x1), x2)
        y3 = sub(mul(lam, sub(x1, x3)), y1)
        return (x3, y3)

    def ec_neg(p1):
        return None if p1 is None else (p1[0], neg(p1[1]))

    def ec_mul(k, p1):
        if k < 0:
            return ec_mul(-k, ec_neg(p1))
        r = None
        while k:
            if k & 1:
                r = ec_add(r, p1)
            p1 = ec_add(p1, p1)
            k >>= 1
        return r

    def on_curve(p1):
        if p1 is None:
            return True
        x, y = p1
        return mul(y, y) == add(mul(mul(x, x), x), b)

    return ec_add, ec_neg, ec_mul, on_curve


g1_add, g1_neg, g1_mul, g1_on_curve = _mk_curve(
    lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P, fp_inv,
    lambda a: -a % P, 0, lambda a: 3 * a % P, B)
g2_add, g2_neg, g2_mul, g2_on_curve = _mk_curve(
    f2_add, f2_sub, f2_mul, f2_inv, f2_neg, F2_ZERO, lambda a: f2_muls(a, 3), f2_muls(XI, B))

G1 = (G1X, G1Y)
G2 = (G2X, G2Y)


def g1_in_subgroup_slow(pt):
    return g1_mul(R, pt) is None


def g1_sigma(pt):
    return None if pt is None else (BETA * pt[0] % P, pt[1])


def g1_in_subgroup_fast(pt):
    """sigma(P) == -[x^2]P  (SURVEY App. A)."""
    if pt is None:
        return True
    return g1_sigma(pt) == g1_neg(g1_mul(X * X, pt))


# ----------------------------------------------------------------------------- serialization (ZCash format)
ST_OK, ST_BAD_FLAGS, ST_X_GE_P, ST_NOT_ON_CURVE, ST_NOT_IN_G1 = 0, 1, 2, 3, 4


def g1_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_decompress(b: bytes, subgroup="fast"):
    """Returns (status, point-or-None).  Validation precedence per SURVEY App. B.2."""
    assert len(b) == 48
    c, inf, sgn = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c:
        return ST_BAD_FLAGS, None
    if inf:
        if sgn or (b[0] & 0x1F) or any(b[1:]):
            return ST_BAD_FLAGS, None
        return ST_OK, None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        return ST_X_GE_P, None
    y = fp_sqrt((x * x * x + B) % P)
    if y is None:
        return ST_NOT_ON_CURVE, None
    if (y > (P - 1) // 2) != bool(sgn):
        y = P - y
    pt = (x, y)
    ok = g1_in_subgroup_fast(pt) if subgroup == "fast" else g1_in_subgroup_slow(pt)
    if not ok:
        return ST_NOT_IN_G1, None
    return ST_OK, pt


def g1_affine_bytes(pt) -> bytes:
    """Canonical 96-byte x||y big-endian; infinity / invalid = 96 zero bytes."""
    if pt is None:
        return bytes(96)
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def _f2_lex_largest(y):
    return y[1] > (P - 1) // 2 if y[1] != 0 else y[0] > (P - 1) // 2


def g2_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(95)
    x, y = pt
    b = bytearray(x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if _f2_lex_largest(y):
        b[0] |= 0x20
    return bytes(b)


def g2_decompress(b: bytes):
    assert len(b) == 96 and b[0] & 0x80 and not b[0] & 0x40
    sgn = b[0] >> 5 & 1
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), f2_muls(XI, B)))
    assert y is not None
    if _f2_lex_largest(y) != bool(sgn):
        y = f2_neg(y)
    return (x, y)


# ----------------------------------------------------------------------------- pairing (textbook)
def _untwist(q):
    """E'(Fp2) -> E(Fp12): (x', y') -> (x'/w^2, y'/w^3)."""
    x, y = q
    w = [F2_ZERO, F2_ONE] + [F2_ZERO] * 4
    w2i = f12_inv(f12_mul(w, w))
    w3i = f12_inv(f12_mul(f12_mul(w, w), w))
    return (f12_mul(f12_from_f2(x), w2i), f12_mul(f12_from_f2(y), w3i))


def miller_loop(p1, q):
    """f_{|x|,Q}(P), conjugated for x < 0.  P in E(Fp) affine, Q in E'(Fp2) affine."""
    if p1 is None or q is None:
        return f12_one()
    qx, qy = _untwist(q)
    px, py = f12_from_f2((p1[0], 0)), f12_from_f2((p1[1], 0))
    three = f12_from_f2((3, 0))
    two = f12_from_f2((2, 0))

    def line(t, s):           # line through t, s evaluated at P; returns (value, t+s)
        tx, ty = t
        sx, sy = s
        if tx == sx and ty == sy:
            lam = f12_mul(f12_mul(three, f12_sqr(tx)), f12_inv(f12_mul(two, ty)))
        else:
            lam = f12_mul(f12_sub(sy, ty), f12_inv(f12_sub(sx, tx)))
        val = f12_sub(f12_sub(py, ty), f12_mul(lam, f12_sub(px, tx)))
        x3 = f12_sub(f12_sub(f12_sqr(lam), tx), sx)
        y3 = f12_sub(f12_mul(lam, f12_sub(tx, x3)), ty)
        return val, (x3, y3)

    f = f12_one()
    t = (qx, qy)
    for i in range(X_ABS.bit_length() - 2, -1, -1):
        l, t2 = line(t, t)
        f = f12_mul(f12_sqr(f), l)
        t = t2
        if X_ABS >> i & 1:
            l, t2 = line(t, (qx, qy))
            f = f12_mul(f, l)
            t = t2
    return f12_conj(f)


FINAL_EXP = (P**12 - 1) // R


def final_exp(f):
    # easy part via Frobenius to keep the pow short, hard part by plain pow
    t = f12_mul(f12_conj(f), f12_inv(f))            # f^(p^6-1)
    t = f12_mul(f12_frob(t, 2), t)                  # ^(p^2+1)
    return f12_pow(t, (P**4 - P**2 + 1) // R)


def pairing(p1, q):
    return final_exp(miller_loop(p1, q))


def pairing_product_is_one(pairs):
    f = f12_one()
    for p1, q in pairs:
        f = f12_mul(f, miller_loop(p1, q))
    return final_exp(f) == f12_one()
