#!/usr/bin/env python3
"""Static per-pipe instruction mix of one kernel, split into its main body and the device functions it calls
(CALL.REL targets).  FMA-pipe cycles per warp and scheduler = 4 x IMAD.WIDE + 2 x other IMAD (measured issue rates,
DESIGN.md 3.1); ALU = IADD3/LOP3/SEL/SHF/PRMT/ISETP/MOV/...  Usage: sass_split.py file.cubin mangled_kernel_name"""
import re
import subprocess
import sys


def mix(lines):
    c = {}
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", l)
        if m:
            c[m.group(1)] = c.get(m.group(1), 0) + 1
    wide = sum(n for op, n in c.items() if op.startswith("IMAD.WIDE"))
    imad = sum(n for op, n in c.items() if op.startswith("IMAD") and not op.startswith("IMAD.WIDE"))
    moved = sum(n for op, n in c.items() if re.match(r"IMAD\.(X|IADD|MOV|SHL)", op))
    alu = sum(n for op, n in c.items() if re.match(r"(IADD3|LOP3|SEL|SHF|PRMT|ISETP|P2R|R2P|MOV|PLOP3|LEA)", op))
    lcl = sum(n for op, n in c.items() if re.match(r"(LDL|STL)", op))
    p2r = sum(n for op, n in c.items() if re.match(r"(P2R|R2P)", op))
    tot = sum(c.values())
    return (f"{tot:5d} instr: IMAD.WIDE {wide:4d}, other IMAD {imad:4d} (X/IADD/MOV/SHL {moved:4d}), ALU {alu:4d} "
            f"(P2R {p2r}), LDL/STL {lcl:3d} => FMA cycles {4 * wide + 2 * imad:5d}, ALU cycles {2 * alu:5d}")


def main():
    cubin, kern = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", kern, cubin], capture_output=True, text=True).stdout.splitlines()
    ins = []
    for l in sass:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), l))
    targets = sorted({int(x, 16) for _, l in ins for x in re.findall(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", l)})
    calls = {}
    for _, l in ins:
        for x in re.findall(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", l):
            calls[int(x, 16)] = calls.get(int(x, 16), 0) + 1
    bounds = targets + [1 << 60]
    first = targets[0] if targets else 1 << 60
    print("main      :", mix([l for a, l in ins if a < first]))
    for i, t in enumerate(targets):
        print(f"fn@{t:#06x} ({calls[t]:2d} call sites):", mix([l for a, l in ins if t <= a < bounds[i + 1]]))


main()
