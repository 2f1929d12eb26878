#!/usr/bin/env python
"""Sum samples by stall reason over the SASS instructions whose execution count lies in
[lo, hi] (e.g. the hot warp's trip).  usage: ncu_sum.py report.ncu-rep lo hi"""
import csv, subprocess, sys, io
rep, lo, hi = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sel = [r for r in data if lo <= int(r[ie]) <= hi]
ts = sum(int(r[sm]) for r in sel); tall = sum(int(r[sm]) for r in data)
print(f"{len(sel)} instrs, samples {ts} of {tall} ({ts/tall:.1%})")
agg = {h: sum(int(r[hdr.index(h)]) for r in sel) for h in stalls}
print("  ".join(f"{h[6:]} {v/ts:.1%}" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
