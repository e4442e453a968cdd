"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: stall samples, executed instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[2] != '-':
        continue
    try:
        smp, ex = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    st = {h: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
    agg[(cur, int(r[0]), r[1])] = (smp, ex, st)
tot = sum(v[0] for v in agg.values())
totex = sum(v[1] for v in agg.values())
print('total samples', tot, 'total warp instructions', totex)
allst = {}
for v in agg.values():
    for k, n in v[2].items():
        allst[k] = allst.get(k, 0) + n
print('stalls:', ', '.join('%s %.1f%%' % (k.replace('stall_', ''), 100.0 * n / max(sum(allst.values()), 1)) for k, n in sorted(allst.items(), key=lambda x: -x[1])[:9]))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    t3 = sorted(v[2].items(), key=lambda x: -x[1])[:3]
    print('%5.1f%% smp %5.1f%% ins  %s:%d  %-72s %s' % (100 * v[0] / tot, 100 * v[1] / totex, k[0], k[1], k[2].strip()[:72],
                                                       ' '.join('%s=%d' % (a.replace('stall_', ''), b) for a, b in t3)))
