#!/usr/bin/env python3
"""Static SASS instruction histogram of the hot kernels: per kernel the opcode counts of the whole
function and of every loop body (backward branch target .. branch), from `cuobjdump -sass` of the built objects.
Proves (a) that every mad.lo.cc / madc.hi.cc pair of the generated Montgomery code became ONE IMAD.WIDE.U32(.X) and
(b) itemises the non-multiply instructions around them.  Usage: python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent.parent / "kzg_batch_verification_scheme_b200" / "csrc"
KERNELS = [("k_decompress.o", "k_decompress_sqrtILi3E"), ("k_msm.o", "k_msm_chunk_pass1_stagedILi4ELi232E"),
           ("k_msm.o", "k_msm_chunk_pass2_t"), ("k_msm.o", "k_red_totals"), ("k_mpair.o", "k_mp_check"), ("k_mpair.o", "k_mp_lines")]
INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", str(CSRC / obj)], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = INS.match(ln)
        if m and cur:
            text = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
            funcs[cur].append((int(m.group(1), 16), text.split()[0], text))
    return funcs


def histo(ins):
    c = collections.Counter(op for _, op, _ in ins)
    return c


def classes(c):
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    imad = sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
    iadd = sum(v for k, v in c.items() if k.startswith("IADD3") or k.startswith("UIADD3"))
    lmem = sum(v for k, v in c.items() if k.startswith(("LDL", "STL")))
    smem = sum(v for k, v in c.items() if k.startswith(("LDS", "STS")))
    gmem = sum(v for k, v in c.items() if k.startswith(("LDG", "STG", "LD.", "ST.", "ATOM", "RED")))
    shfl = sum(v for k, v in c.items() if k.startswith("SHFL"))
    tot = sum(c.values())
    return dict(total=tot, imad_wide=wide, imad_other=imad, iadd3=iadd, local_mem=lmem, shared_mem=smem, global_mem=gmem, shfl=shfl,
                other=tot - wide - imad - iadd - lmem - smem - gmem - shfl)


def main():
    cache = {}
    for obj, key in KERNELS:
        funcs = cache.setdefault(obj, functions(obj))
        names = [n for n in funcs if key in n]
        if not names:
            print(f"## {key}: not found in {obj}")
            continue
        ins = funcs[names[0]]
        c = histo(ins)
        cl = classes(c)
        print(f"## {names[0]}  ({obj})")
        print("   whole function (static): " + ", ".join(f"{k} {v}" for k, v in cl.items()) +
              f"   -> IMAD.WIDE share {cl['imad_wide'] / max(cl['total'], 1):.3f}")
        print("   top opcodes: " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
        # loops: backward branches
        addr = [a for a, _, _ in ins]
        loops = []
        for a, op, text in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if m and int(m.group(1), 16) < a:
                    loops.append((int(m.group(1), 16), a))
        for lo, hi in sorted(loops):
            body = [x for x in ins if lo <= x[0] <= hi]
            cb = classes(histo(body))
            print(f"   loop 0x{lo:x}..0x{hi:x} (static body): " + ", ".join(f"{k} {v}" for k, v in cb.items()) +
                  f"   -> IMAD.WIDE share {cb['imad_wide'] / max(cb['total'], 1):.3f}")
        mad_pairs = sum(1 for _, op, _ in ins if op in ("IMAD.HI.U32",))
        print(f"   unfused multiply halves (IMAD.HI.U32, IMAD.LO paired with IADD3 carries): IMAD.HI.U32 {mad_pairs}")
        print()


if __name__ == "__main__":
    sys.exit(main())
