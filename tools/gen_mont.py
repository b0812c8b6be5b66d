#!/usr/bin/env python3
"""Generates kzg_batch_verification_scheme_b200/csrc/mont_gen.cuh: Montgomery multiplication / squaring
and add/sub carry chains for Fp (12 x 32-bit limbs) and Fr (8 x 32-bit limbs) as single inline-PTX
blocks (BASELINE.json:5 item (a): "Fp/Fr Montgomery arithmetic in 12x32-bit and 8x32-bit limbs with
carry chains in inline PTX").

Every routine is first built as a tiny instruction list (IR); the same list is (1) printed as PTX and
(2) executed by a Python interpreter with an explicit carry flag (`simulate`), so `--selftest` proves
the emitted instruction stream against big-int arithmetic without a GPU.

Multiplication scheme (even/odd accumulators): the 64-bit products a_j*b_i for even j touch limb pairs
(j, j+1) that do not overlap, so they can be accumulated into one array with ONE carry chain; odd j go
to a second array that is offset by one limb.  Each mad.lo.cc/madc.hi.cc pair fuses to one
IMAD.WIDE.U32 in SASS, giving 2N^2+N wide multiply-adds per product.  After each b_i the reduction
m = T[0]*(-p^-1) is folded in the same way and the roles of the two arrays swap (the shift by one
limb changes their parity).
"""
from __future__ import annotations

import argparse
import random
import sys
from pathlib import Path

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
MASK = 0xFFFFFFFF


def limbs(v, n):
    return [(v >> (32 * i)) & MASK for i in range(n)]


# ----------------------------------------------------------------------------- IR + interpreter
class Prog:
    def __init__(self):
        self.ins = []

    def op(self, name, d, *src):
        self.ins.append((name, d, src))


def simulate(prog: Prog, env: dict) -> dict:
    """env: register name -> u32.  Immediates are ints.  Returns env (mutated)."""
    cf = 0

    def val(x):
        return x if isinstance(x, int) else env[x]

    for name, d, src in prog.ins:
        s = [val(x) for x in src]
        base = name.replace(".cc", "").replace(".u32", "")
        cin = 0
        if base in ("madc.lo", "madc.hi", "addc", "subc"):
            cin = cf
        if base in ("mul.lo",):
            t = (s[0] * s[1]) & MASK
        elif base in ("mul.hi",):
            t = (s[0] * s[1]) >> 32
        elif base in ("mad.lo", "madc.lo"):
            t = ((s[0] * s[1]) & MASK) + s[2] + cin
        elif base in ("mad.hi", "madc.hi"):
            t = ((s[0] * s[1]) >> 32) + s[2] + cin
        elif base in ("add", "addc"):
            t = s[0] + s[1] + cin
        elif base in ("sub", "subc"):
            t = s[0] - s[1] - cin
        elif base == "mov":
            t = s[0]
        else:
            raise ValueError(name)
        if ".cc" in name:
            cf = 1 if (t >> 32) & 1 or t < 0 else 0
            if base in ("sub", "subc"):
                cf = 1 if t < 0 else 0
        env[d] = t & MASK
    return env


# ----------------------------------------------------------------------------- routines
def cmad_chain(pg, acc, a_regs, b, first_has_carry_in=False, last_cc=True):
    """acc[2k],acc[2k+1] += a_regs[k]*b with one carry chain."""
    for k, a in enumerate(a_regs):
        lo = "mad.lo.cc.u32" if (k == 0 and not first_has_carry_in) else "madc.lo.cc.u32"
        pg.op(lo, acc[2 * k], a, b, acc[2 * k])
        last = k == len(a_regs) - 1
        pg.op("madc.hi.cc.u32" if (not last or last_cc) else "madc.hi.u32", acc[2 * k + 1], a, b, acc[2 * k + 1])


def gen_mul(n, mod, a, b, out, sqr=False):
    """Montgomery product a*b/2^(32n) mod `mod`, result in [0, 2*mod) written to regs `out`.
    a, b, out: lists of register names.  Internal regs e*, o*, m."""
    pl = limbs(mod, n)
    m0 = (-pow(mod, -1, 1 << 32)) & MASK
    pg = Prog()
    E = [f"e{k}" for k in range(n)]
    O = [f"o{k}" for k in range(n)]
    for i in range(n):
        X, Y = (E, O) if i % 2 == 0 else (O, E)      # X[k] at limb k ; Y[k] at limb k-1 (Y[0] dead)
        bi = b[i]
        if i == 0:
            for k in range(0, n, 2):
                pg.op("mul.lo.u32", Y[k], a[k + 1], bi)
                pg.op("mul.hi.u32", Y[k + 1], a[k + 1], bi)
            for k in range(0, n, 2):
                pg.op("mul.lo.u32", X[k], a[k], bi)
                pg.op("mul.hi.u32", X[k + 1], a[k], bi)
        else:
            pg.op("add.cc.u32", X[0], X[0], Y[1])
            for k in range(0, n - 2, 2):             # Y'[k] = lo(a[k+1] bi) + Y[k+2] (+carry)
                pg.op("madc.lo.cc.u32", Y[k], a[k + 1], bi, Y[k + 2])
                pg.op("madc.hi.cc.u32", Y[k + 1], a[k + 1], bi, Y[k + 3])
            pg.op("madc.lo.cc.u32", Y[n - 2], a[n - 1], bi, 0)
            pg.op("madc.hi.u32", Y[n - 1], a[n - 1], bi, 0)
            cmad_chain(pg, X, [a[k] for k in range(0, n, 2)], bi)
            pg.op("addc.u32", Y[n - 1], Y[n - 1], 0)
        pg.op("mul.lo.u32", "m", X[0], m0)
        cmad_chain(pg, Y, [pl[k] for k in range(1, n, 2)], "m", last_cc=False)
        cmad_chain(pg, X, [pl[k] for k in range(0, n, 2)], "m")
        pg.op("addc.u32", Y[n - 1], Y[n - 1], 0)
    # n even: last iteration used (X,Y) = (O,E): result[k] = E[k] + O[k+1]
    X, Y = (E, O) if (n - 1) % 2 == 0 else (O, E)
    for k in range(n - 1):
        pg.op("add.cc.u32" if k == 0 else "addc.cc.u32", out[k], Y[k], X[k + 1])
    pg.op("addc.u32", out[n - 1], Y[n - 1], 0)
    return pg, E + O + ["m"]


def gen_sqr(n, mod, a, out):
    """Montgomery square a*a/2^(32n), result in [0, 2*mod).  n(n-1)/2 + n wide multiply-adds for the
    2n-limb square (off-diagonal products once, doubled, plus the diagonal) and n^2 + n for the reduction
    of its low half ("multiply by 1" with the limb shift fused into the m*p_odd chain): 234 for n = 12.

    Off-diagonal products a_i*a_(i+d) sit at limbs (2i+d, 2i+d+1): for a fixed difference d they do not
    overlap, so each d is ONE carry chain; even d go to array ev*, odd d to od* (true limb positions).
    Processing d from large to small makes every chain's carry-out land in a limb no earlier chain has
    touched."""
    pl = limbs(mod, n)
    m0 = (-pow(mod, -1, 1 << 32)) & MASK
    pg = Prog()
    EV = [f"v{k}" for k in range(2 * n)]
    OD = [f"w{k}" for k in range(2 * n)]
    for arr, ds in ((EV, range(n - 2, 0, -2)), (OD, range(n - 1, 0, -2))):
        touched = set()
        for d in ds:
            cnt = n - d
            chain_open = False
            for i in range(cnt):
                lo, hi = 2 * i + d, 2 * i + d + 1
                alo = arr[lo] if lo in touched else 0
                ahi = arr[hi] if hi in touched else 0
                if not chain_open and alo == 0 and ahi == 0:
                    pg.op("mul.lo.u32", arr[lo], a[i], a[i + d])          # fresh pair, nothing to carry
                    pg.op("mul.hi.u32", arr[hi], a[i], a[i + d])
                else:
                    pg.op("madc.lo.cc.u32" if chain_open else "mad.lo.cc.u32", arr[lo], a[i], a[i + d], alo)
                    pg.op("madc.hi.cc.u32", arr[hi], a[i], a[i + d], ahi)
                    chain_open = True
                touched.update((lo, hi))
            if chain_open:
                top = 2 * (cnt - 1) + d + 2
                assert top not in touched
                pg.op("addc.u32", arr[top], 0, 0)
                touched.add(top)
        for k in range(2 * n):
            if k not in touched:
                pg.op("mov.u32", arr[k], 0)
    # S = EV + OD ; T = 2S + diag.  (S < 2^(64n-1), so nothing is lost)
    T = [f"t{k}" for k in range(2 * n)]
    for k in range(2 * n):
        pg.op("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < 2 * n - 1 else "addc.u32"), T[k], EV[k], OD[k])
    for k in range(2 * n):
        pg.op("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < 2 * n - 1 else "addc.u32"), T[k], T[k], T[k])
    for i in range(n):
        pg.op("mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32", T[2 * i], a[i], a[i], T[2 * i])
        pg.op("madc.hi.cc.u32" if i < n - 1 else "madc.hi.u32", T[2 * i + 1], a[i], a[i], T[2 * i + 1])
    # reduction of the low half: X = T[0..n) at limb k, Y = 0 at limb k+1; roles swap every iteration
    A = T[:n]
    Bq = [f"q{k}" for k in range(n)]
    for i in range(n):
        X, Y = (A, Bq) if i % 2 == 0 else (Bq, A)
        if i == 0:
            pg.op("mul.lo.u32", "m", X[0], m0)
            for k in range(0, n, 2):
                pg.op("mul.lo.u32", Y[k], "m", pl[k + 1])
                pg.op("mul.hi.u32", Y[k + 1], "m", pl[k + 1])
        else:
            pg.op("add.cc.u32", X[0], X[0], Y[1])
            pg.op("mul.lo.u32", "m", X[0], m0)                      # does not touch the carry flag
            for k in range(0, n - 2, 2):
                pg.op("madc.lo.cc.u32", Y[k], "m", pl[k + 1], Y[k + 2])
                pg.op("madc.hi.cc.u32", Y[k + 1], "m", pl[k + 1], Y[k + 3])
            pg.op("madc.lo.cc.u32", Y[n - 2], "m", pl[n - 1], 0)
            pg.op("madc.hi.u32", Y[n - 1], "m", pl[n - 1], 0)
        cmad_chain(pg, X, [pl[k] for k in range(0, n, 2)], "m")
        pg.op("addc.u32", Y[n - 1], Y[n - 1], 0)
    X, Y = (A, Bq) if (n - 1) % 2 == 0 else (Bq, A)
    U = [f"u{k}" for k in range(n)]
    for k in range(n - 1):
        pg.op("add.cc.u32" if k == 0 else "addc.cc.u32", U[k], Y[k], X[k + 1])
    pg.op("addc.u32", U[n - 1], Y[n - 1], 0)
    for k in range(n):
        pg.op("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < n - 1 else "addc.u32"), out[k], U[k], T[n + k])
    return pg, EV + OD + T + Bq + U + ["m"]


def check_sqr(n, mod, trials=300):
    rnd = random.Random(50 + n)
    a = [f"a{k}" for k in range(n)]
    out = [f"r{k}" for k in range(n)]
    pg, _ = gen_sqr(n, mod, a, out)
    rinv = pow(1 << (32 * n), -1, mod)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << (32 * n)) % mod, mod >> 1, (1 << (32 * n - 3)) - 1,
            sum(0xFFFFFFFF << (64 * k) for k in range(n // 2)) % mod]
    lazy = 4 * mod < (1 << (32 * n))
    for x in edge + ([mod, mod + 1, 2 * mod - 1] if lazy else []) + [rnd.randrange(2 * mod if lazy else mod) for _ in range(trials)]:
        env = {a[k]: limbs(x, n)[k] for k in range(n)}
        simulate(pg, env)
        got = sum(env[out[k]] << (32 * k) for k in range(n))
        assert got < 2 * mod and got % mod == x * x * rinv % mod, (n, hex(x))
    nim = sum(1 for nm, _, _ in pg.ins if ".hi" in nm) + sum(1 for nm, _, _ in pg.ins if nm == "mul.lo.u32" and False)
    return len(pg.ins), nim


def emit_sqr_fn(fname, n, mod):
    a = [f"a{k}" for k in range(n)]
    out = [f"r{k}" for k in range(n)]
    pg, temps = gen_sqr(n, mod, a, out)
    body = emit_asm(pg, [(out[k], f"r[{k}]") for k in range(n)], [(a[k], f"a[{k}]") for k in range(n)], temps)
    nim = sum(1 for nm, _, _ in pg.ins if ".hi" in nm)
    return (f"// Montgomery square, result in [0, 2p).  {nim} wide multiply-adds + {n} mul.lo.\n"
            f"__device__ __forceinline__ void {fname}(uint32_t (&r)[{n}], const uint32_t (&a)[{n}]) {{\n"
            f"{body}\n}}\n")


def gen_add(n, mod):
    """r = a + b (no overflow: a,b < p < 2^(32n-1)); t = r - p; bw = all-ones iff r < p."""
    pl = limbs(mod, n)
    pg = Prog()
    for k in range(n):
        pg.op("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < n - 1 else "addc.u32"), f"r{k}", f"a{k}", f"b{k}")
    for k in range(n):
        pg.op("sub.cc.u32" if k == 0 else "subc.cc.u32", f"t{k}", f"r{k}", pl[k])
    pg.op("subc.u32", "bw", 0, 0)
    return pg


def gen_sub(n, mod):
    """r = a - b mod 2^(32n); bw = all-ones iff a < b; t = r + p."""
    pl = limbs(mod, n)
    pg = Prog()
    for k in range(n):
        pg.op("sub.cc.u32" if k == 0 else "subc.cc.u32", f"r{k}", f"a{k}", f"b{k}")
    pg.op("subc.u32", "bw", 0, 0)
    for k in range(n):
        pg.op("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < n - 1 else "addc.u32"), f"t{k}", f"r{k}", pl[k])
    return pg


def gen_reduce(n, mod):
    """t = a - p; bw = all-ones iff a < p."""
    pl = limbs(mod, n)
    pg = Prog()
    for k in range(n):
        pg.op("sub.cc.u32" if k == 0 else "subc.cc.u32", f"t{k}", f"a{k}", pl[k])
    pg.op("subc.u32", "bw", 0, 0)
    return pg


# ----------------------------------------------------------------------------- PTX emission
def emit_asm(pg: Prog, outs, ins, temps, indent="    "):
    """outs/ins: lists of (reg name in IR, C expression).  Returns C++ asm statement text."""
    opmap = {}
    idx = 0
    for name, _ in outs:
        opmap[name] = f"%{idx}"
        idx += 1
    for name, _ in ins:
        opmap[name] = f"%{idx}"
        idx += 1

    def fmt(x):
        if isinstance(x, int):
            return f"0x{x:08x}"
        return opmap.get(x, x)

    lines = [f'{indent}asm("{{\\n\\t"']
    temps = list(temps) + ["zr"]
    lines.append(f'{indent}    ".reg .u32 {", ".join(temps)};\\n\\t"')
    lines.append(f'{indent}    "mov.u32 zr, 0;\\n\\t"')
    for name, d, src in pg.ins:
        src = list(src)
        if name.startswith(("mad", "mul")) and isinstance(src[0], int):
            src[0], src[1] = src[1], src[0]           # register operand first, immediate second
        if name.startswith(("add", "sub")) and isinstance(src[0], int):
            assert src[0] == 0
            src[0] = "zr"
        lines.append(f'{indent}    "{name} {fmt(d)}, {", ".join(fmt(s) for s in src)};\\n\\t"')
    lines.append(f'{indent}    "}}"')
    lines.append(f'{indent}    : {", ".join(f"""\"=&r\"({c})""" for _, c in outs)}')
    lines.append(f'{indent}    : {", ".join(f"""\"r\"({c})""" for _, c in ins)});')
    return "\n".join(lines)


def emit_mul_fn(fname, n, mod):
    a = [f"a{k}" for k in range(n)]
    b = [f"b{k}" for k in range(n)]
    out = [f"r{k}" for k in range(n)]
    pg, temps = gen_mul(n, mod, a, b, out)
    body = emit_asm(pg, [(out[k], f"r[{k}]") for k in range(n)],
                    [(a[k], f"a[{k}]") for k in range(n)] + [(b[k], f"b[{k}]") for k in range(n)], temps)
    nim = sum(1 for nm, _, _ in pg.ins if ".hi" in nm)
    return (f"// Montgomery product, result in [0, 2p).  {nim} wide multiply-adds + {n} mul.lo.\n"
            f"__device__ __forceinline__ void {fname}(uint32_t (&r)[{n}], const uint32_t (&a)[{n}], const uint32_t (&b)[{n}]) {{\n"
            f"{body}\n}}\n")


def emit_addsub_fns(prefix, n, mod):
    out = []
    a_in = [(f"a{k}", f"a[{k}]") for k in range(n)]
    b_in = [(f"b{k}", f"b[{k}]") for k in range(n)]
    r_out = [(f"r{k}", f"r[{k}]") for k in range(n)]
    t_out = [(f"t{k}", f"t[{k}]") for k in range(n)]
    sel_lt = "    #pragma unroll\n    for (int k = 0; k < %d; ++k) r[k] = bw ? r[k] : t[k];\n" % n
    sel_ge = "    #pragma unroll\n    for (int k = 0; k < %d; ++k) r[k] = bw ? t[k] : r[k];\n" % n
    body = emit_asm(gen_add(n, mod), r_out + t_out + [("bw", "bw")], a_in + b_in, [])
    out.append(f"// r = a + b mod p (inputs and output canonical)\n"
               f"__device__ __forceinline__ void {prefix}_add_ptx(uint32_t (&r)[{n}], const uint32_t (&a)[{n}], const uint32_t (&b)[{n}]) {{\n"
               f"    uint32_t t[{n}], bw;\n{body}\n{sel_lt}}}\n")
    body = emit_asm(gen_sub(n, mod), r_out + t_out + [("bw", "bw")], a_in + b_in, [])
    out.append(f"// r = a - b mod p\n"
               f"__device__ __forceinline__ void {prefix}_sub_ptx(uint32_t (&r)[{n}], const uint32_t (&a)[{n}], const uint32_t (&b)[{n}]) {{\n"
               f"    uint32_t t[{n}], bw;\n{body}\n{sel_ge}}}\n")
    body = emit_asm(gen_reduce(n, mod), t_out + [("bw", "bw")], [(f"a{k}", f"r[{k}]") for k in range(n)], [])
    out.append(f"// r in [0,2p) -> [0,p)\n"
               f"__device__ __forceinline__ void {prefix}_reduce_ptx(uint32_t (&r)[{n}]) {{\n"
               f"    uint32_t t[{n}], bw;\n{body}\n{sel_lt}}}\n")
    return "\n".join(out)


def check_addsub_lazy(n, mod, trials=300):
    """add/sub generated for the modulus 2p, operands in [0, 2p): results in [0, 2p) and correct mod p."""
    rnd = random.Random(200 + n)
    m2 = 2 * mod
    edge = [0, 1, mod - 1, mod, mod + 1, m2 - 1, m2 - 2]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(m2), rnd.randrange(m2)) for _ in range(trials)]
    for x, y in cases:
        env = {}
        for k in range(n):
            env[f"a{k}"] = limbs(x, n)[k]
            env[f"b{k}"] = limbs(y, n)[k]
        e = simulate(gen_add(n, m2), dict(env))
        r = sum(e[f"r{k}"] << (32 * k) for k in range(n)); t = sum(e[f"t{k}"] << (32 * k) for k in range(n))
        got = r if e["bw"] else t
        assert got < m2 and got % mod == (x + y) % mod
        e = simulate(gen_sub(n, m2), dict(env))
        r = sum(e[f"r{k}"] << (32 * k) for k in range(n)); t = sum(e[f"t{k}"] << (32 * k) for k in range(n))
        got = t if e["bw"] else r
        assert got < m2 and got % mod == (x - y) % mod


def check_addsub(n, mod, trials=300):
    rnd = random.Random(100 + n)
    edge = [0, 1, mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(trials)]
    for x, y in cases:
        env = {}
        for k in range(n):
            env[f"a{k}"] = limbs(x, n)[k]
            env[f"b{k}"] = limbs(y, n)[k]
        e = simulate(gen_add(n, mod), dict(env))
        r = sum(e[f"r{k}"] << (32 * k) for k in range(n)); t = sum(e[f"t{k}"] << (32 * k) for k in range(n))
        assert (r if e["bw"] else t) == (x + y) % mod and e["bw"] in (0, MASK)
        e = simulate(gen_sub(n, mod), dict(env))
        r = sum(e[f"r{k}"] << (32 * k) for k in range(n)); t = sum(e[f"t{k}"] << (32 * k) for k in range(n))
        assert (t if e["bw"] else r) == (x - y) % mod
        v = x + y                                    # < 2p
        env2 = {f"a{k}": limbs(v, n)[k] for k in range(n)}
        e = simulate(gen_reduce(n, mod), env2)
        t = sum(e[f"t{k}"] << (32 * k) for k in range(n))
        assert (v if e["bw"] else t) == v % mod


def check_mul(n, mod, trials=300):
    rnd = random.Random(n)
    a = [f"a{k}" for k in range(n)]
    b = [f"b{k}" for k in range(n)]
    out = [f"r{k}" for k in range(n)]
    pg, _ = gen_mul(n, mod, a, b, out)
    rinv = pow(1 << (32 * n), -1, mod)
    lazy = 4 * mod < (1 << (32 * n))          # operands may live in [0, 2p) when the modulus has >= 2 spare bits
    top = 2 * mod if lazy else mod
    edge = [0, 1, mod - 1, mod - 2, (1 << (32 * n)) % mod, mod >> 1] + ([mod, mod + 1, 2 * mod - 1] if lazy else [])
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(top), rnd.randrange(top)) for _ in range(trials)]
    for x, y in cases:
        env = {}
        for k in range(n):
            env[a[k]] = limbs(x, n)[k]
            env[b[k]] = limbs(y, n)[k]
        simulate(pg, env)
        got = sum(env[out[k]] << (32 * k) for k in range(n))
        assert got < 2 * mod and got % mod == x * y * rinv % mod, (n, hex(x), hex(y))
    return len(pg.ins)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--selftest", action="store_true")
    ap.add_argument("-o", default=str(Path(__file__).resolve().parent.parent / "kzg_batch_verification_scheme_b200" / "csrc" / "mont_gen.cuh"))
    args = ap.parse_args()
    n1 = check_mul(12, P)
    n2 = check_mul(8, R)
    check_addsub(12, P)
    check_addsub(8, R)
    check_addsub_lazy(12, P)
    ns, nim = check_sqr(12, P)
    check_sqr(8, R)
    if args.selftest:
        print(f"selftest ok: fp_mul {n1} instrs, fr_mul {n2} instrs, fp_sqr {ns} instrs ({nim} wide mads)")
        return
    txt = ["// GENERATED by tools/gen_mont.py -- do not edit.  Inline-PTX Montgomery products (sm_100a).",
           "#pragma once", "#include <cstdint>", "",
           emit_mul_fn("fp_mont_mul_ptx", 12, P), emit_mul_fn("fr_mont_mul_ptx", 8, R), emit_sqr_fn("fp_mont_sqr_ptx", 12, P),
           emit_addsub_fns("fp", 12, P), emit_addsub_fns("fr", 8, R),
           "// lazy domain: operands and results in [0, 2p) (modulus 2p)\n" + emit_addsub_fns("fp2p", 12, 2 * P)]
    Path(args.o).write_text("\n".join(txt))
    print(f"wrote {args.o}: fp_mul {n1} instrs, fr_mul {n2} instrs")


if __name__ == "__main__":
    main()
