#!/usr/bin/env python
"""Print the SASS instructions of an ncu report whose execution count lies in [lo, hi], with
samples and the main stall reasons.  usage: ncu_range.py report.ncu-rep lo hi"""
import csv, subprocess, sys, io
rep, lo, hi = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
names = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_math", "stall_lg", "stall_not_selected", "stall_barrier", "stall_dispatch"]
idx = [hdr.index(n) for n in names if n in hdr]
print("idx exec samples " + " ".join(n[6:] for n in names if n in hdr))
tot = 0; n_i = 0
for i, r in enumerate(data):
    n = int(r[ie])
    if lo <= n <= hi:
        tot += int(r[sm]); n_i += 1
        print(f"{i:5d} {n:9d} {int(r[sm]):6d} " + " ".join(f"{int(r[j]):5d}" for j in idx) + "  " + r[1].strip()[:100])
print(f"{n_i} instructions, {tot} samples of {sum(int(r[sm]) for r in data)}")
