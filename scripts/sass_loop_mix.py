#!/usr/bin/env python3
"""Instruction mix of the innermost loops of one kernel (cuobjdump -sass of an object file): for every backward branch,
the opcode histogram of the address range it closes. Usage: sass_loop_mix.py <object> <mangled-function-substring>"""
import collections
import re
import subprocess
import sys


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    blocks = txt.split("Function : ")
    body = next(b for b in blocks if b.split("\n", 1)[0].strip().find(fun) >= 0)
    ins = []
    for ln in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(body.split("\n", 1)[0].strip(), len(ins), "instructions")
    loops = []
    for a, t in ins:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in loops:
        inner = [t for a, t in ins if lo <= a <= hi]
        h = collections.Counter()
        for t in inner:
            t = re.sub(r"^@!?U?P\d\s+", "", t)
            h[t.split()[0].split(".")[0]] += 1
        if len(inner) < 40:
            continue
        print(f"loop {lo:#x}..{hi:#x}: {len(inner)} instr;", ", ".join(f"{k} {v}" for k, v in h.most_common(14)))


if __name__ == "__main__":
    main()
