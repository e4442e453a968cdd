"""Aggregates an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and stall samples."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = defaultdict(lambda: [0, 0])
stalls = defaultdict(int)
tot = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex]:
        continue
    op = r[isrc].strip().split()
    if not op:
        continue
    name = op[1] if op[0].startswith("@") else op[0]
    name = name.split(".")[0] + ("." + ".".join(name.split(".")[1:2]) if name.startswith(("MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "LDL", "STL")) and "." in name else "")
    n = int(r[iex])
    ops[name][0] += n
    ops[name][1] += int(r[ismp] or 0)
    tot += n
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print("total warp instructions %d  (%.0f per env step)  static instructions %d" % (tot, tot / steps, len(rows) - 2))
for k, (n, s) in sorted(ops.items(), key=lambda x: -x[1][0])[:40]:
    print("%-14s %14d %6.2f%%  per-step %8.1f  samples %d" % (k, n, 100.0 * n / tot, n / steps, s))
ts = sum(stalls.values())
print("stall samples:", ", ".join("%s %.1f%%" % (k, 100.0 * v / ts) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:10]))
