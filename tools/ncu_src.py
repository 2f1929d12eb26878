#!/usr/bin/env python
"""Per-region / per-instruction stall samples of one kernel table of `ncu --page source --csv --print-source sass`.
usage: ncu_src.py src.csv TABLE_INDEX [lo hi]   (TABLE_INDEX: 0-based among the 'Kernel Name' tables)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
ti = int(sys.argv[2])
lo_r = starts[ti]; hi_r = starts[ti + 1] if ti + 1 < len(starts) else len(rows)
hdr = rows[lo_r + 1]
data = [r for r in rows[lo_r + 2:hi_r] if len(r) == len(hdr)]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); src = hdr.index("Source")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[sm]) for r in data)
print(rows[lo_r][1], "instructions", len(data), "samples", tot)
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    agg = collections.Counter()
    for i in range(lo, hi):
        r = data[i]
        for n in names: agg[n[6:]] += int(r[hdr.index(n)])
    s = sum(int(data[i][sm]) for i in range(lo, hi))
    print(f"region {lo}-{hi}: samples {s} ({s/tot:.1%})", {k: v for k, v in agg.most_common(8)})
    thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
    for i in range(lo, hi):
        r = data[i]
        if int(r[sm]) > thr * s:
            top = sorted(((int(r[hdr.index(n)]), n[6:]) for n in names), reverse=True)[:2]
            print(i, r[ie], r[sm], top, r[src].strip()[:70])
else:
    reg = collections.OrderedDict()
    for i, r in enumerate(data):
        a = reg.setdefault(i // 50, [0, 0]); a[0] += int(r[sm]); a[1] = max(a[1], int(r[ie]))
    for k, (s, n) in reg.items():
        if s > tot * 0.005: print(f"{k*50:5d}-{k*50+49:5d} samples {s:7d} {s/tot:6.1%}  max exec {n:12d}")
