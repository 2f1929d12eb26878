#!/usr/bin/env python
"""Per-instruction stall table for a SASS index range of an ncu report.
usage: ncu_regions.py report.ncu-rep lo hi [min_samples] [symbols_per_stream] [warps]"""
import csv, subprocess, io, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
sym = int(sys.argv[5]) if len(sys.argv) > 5 else 43425
warps = int(sys.argv[6]) if len(sys.argv) > 6 else 1024
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
names = [n for n in ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_sleeping", "stall_lg", "stall_mio", "stall_not_selected", "stall_barrier", "stall_membar"] if n in hdr]
idx = [hdr.index(n) for n in names]
print("idx exec/step samples | " + " ".join(n[6:12] for n in names))
step = warps * sym; tot = 0; allc = sum(int(r[sm]) for r in data); ninst = 0
for i in range(lo, min(hi, len(data))):
    r = data[i]; tot += int(r[sm]); ninst += int(r[ie])
    if int(r[sm]) >= mins:
        print(f"{i:5d} {int(r[ie])/step:5.2f} {int(r[sm]):7d} " + " ".join(f"{int(r[j]):6d}" for j in idx) + "  " + r[1].strip()[:80])
print(f"region: {tot} samples = {tot/allc:.1%} of kernel; {ninst/step:.1f} warp-instr per step")
