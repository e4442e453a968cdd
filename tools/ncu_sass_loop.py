"""Per-SASS-instruction warp-state samples of the hottest MUFU loop of a kernel.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > src_sass.csv ; python tools/ncu_sass_loop.py src_sass.csv
"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
data = rows[2:]
tot = sum(int(r[ismp]) for r in data)
best = max(range(len(data)), key=lambda i: int(data[i][iex]) if 'MUFU.EX2' in data[i][ia] else 0)
lo = best
while lo > 0 and 'BRA' not in data[lo][ia] and 'BSSY' not in data[lo][ia]:
    lo -= 1
hi = best
while 'BRA' not in data[hi][ia]:
    hi += 1
print('loop', lo, hi, 'samples share %.3f' % (sum(int(r[ismp]) for r in data[lo:hi + 1]) / tot), 'total samples', tot)
for r in data[lo:hi + 1]:
    s, ex = int(r[ismp]), int(r[iex])
    top = sorted(((int(r[i] or 0), h.replace('stall_', '')) for i, h in st), reverse=True)[:3]
    print("%-64s smp %6d  ex %10d  %5.2f  %s" % (r[ia].strip()[:64], s, ex, s / (ex / 1e6 + 1e-9), ' '.join('%s=%d' % (h, n) for n, h in top if n)))
