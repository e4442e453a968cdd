#!/usr/bin/env python
"""Static SASS view of a kernel's hottest loop (no GPU needed).

    python tools/sass_loop.py <object-or-so> <mangled kernel name> [--seq]

Finds every backward branch, picks the innermost loop with the most MUFU instructions (the TD chunk loop of the fused
kernel), prints its opcode histogram and, with --seq, the issue-order string (M = MUFU, F = packed/scalar FP on the FMA
pipe, A = ALU-pipe op, L = LDS/STS/LDG, S = SHFL, . = other) so that MUFU clustering is visible.
"""
import collections
import re
import subprocess
import sys


def load(obj, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    ins = []
    for l in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    return ins


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0]


def klass(op):
    b = op.split(".")[0]
    if b == "MUFU": return "M"
    if b in ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "HFMA2", "DFMA", "DMUL", "DADD"): return "F"
    if b in ("FSEL", "ISETP", "FSETP", "LOP3", "IADD3", "IMAD", "LEA", "MOV", "SEL", "FMNMX", "VIADD", "PLOP3", "SHF", "IMNMX", "VIMNMX", "PRMT", "UMOV", "CS2R"): return "A"
    if b in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDL", "STL", "LDC", "LDSM"): return "L"
    if b == "SHFL": return "S"
    return "."


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    ins = load(obj, fun)
    a2i = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in a2i:
                j = a2i[int(m.group(1), 16)]
                mu = sum("MUFU" in s for _, s in ins[j:i + 1])
                loops.append((i - j + 1, mu, j, i))
    # the TD chunk loop: the smallest loop that holds (nearly) the largest number of packed FMAs of any loop <= 3000 instructions
    def ffma2(j, i):
        return sum(opcode(t).startswith("FFMA2") for _, t in ins[j:i + 1])
    cand = [(n, ffma2(j, i), j, i) for n, mu, j, i in loops if n <= 3000]
    top = max(c[1] for c in cand)
    best = min((c for c in cand if c[1] >= 0.8 * top), key=lambda c: c[0])
    best = (best[0], sum("MUFU" in t for _, t in ins[best[2]:best[3] + 1]), best[2], best[3])
    n, mu, j, i = best
    body = ins[j:i + 1]
    print("kernel instructions %d; hottest loop %#x..%#x: %d instructions, %d MUFU" % (len(ins), ins[j][0], ins[i][0], n, mu))
    h = collections.Counter(opcode(t) for _, t in body)
    for op, c in h.most_common():
        print("  %-14s %4d" % (op, c))
    kc = collections.Counter(klass(opcode(t)) for _, t in body)
    print("classes:", dict(kc))
    if "--seq" in sys.argv:
        s = "".join(klass(opcode(t)) for _, t in body)
        for k in range(0, len(s), 100):
            print(s[k:k + 100])


if __name__ == "__main__":
    main()
