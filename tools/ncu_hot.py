#!/usr/bin/env python
"""Summarise an ncu report of the inflate kernel: stall mix, the hot loop's per-instruction
samples and the share of samples outside the loop.  usage: ncu_hot.py report.ncu-rep [symbols_per_stream] [warps] [-v]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
sym = int(sys.argv[2]) if len(sys.argv) > 2 else 43425
warps = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
verbose = "-v" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:pz_inflate"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data); ts = sum(int(r[sm]) for r in data)
step = warps * sym
print(f"warp-instructions {tot}  = {tot/step:.1f} per warp-step; samples {ts}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[hdr.index(h)]) for r in data) for h in stalls}
print("  ".join(f"{h[6:]} {v/ts:.1%}" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
names = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_math", "stall_lg"]
idx = [hdr.index(n) for n in names]
loop = 0; nloop = 0
for i, r in enumerate(data):
    n = int(r[ie])
    if n > 0.25 * step:
        loop += int(r[sm]); nloop += n
        if verbose:
            print(f"{i:5d} {n/step:5.2f} {int(r[sm]):7d} " + " ".join(f"{int(r[j]):6d}" for j in idx) + "  " + r[1].strip()[:90])
print(f"hot loop: {nloop/step:.1f} instr per warp-step, {loop/ts:.1%} of samples")
reg = collections.OrderedDict()
for i, r in enumerate(data):
    n = int(r[ie])
    if n > 0.25 * step: continue
    a = reg.setdefault(i // 100, [0, 0]); a[0] += int(r[sm]); a[1] += n
for k, (s, n) in sorted(reg.items(), key=lambda x: -x[1][0])[:8]:
    print(f"  outside: instr {k*100:5d}-{k*100+99:5d}: samples {s/ts:.2%} exec/step {n/step:6.2f}")
