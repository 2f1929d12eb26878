#!/usr/bin/env python
"""Group the SASS instructions of an ncu report by execution count (instructions of one loop
share their count) and list the groups by samples.  usage: ncu_groups.py report.ncu-rep [top]"""
import csv, subprocess, io, collections, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
c = collections.defaultdict(lambda: [0, 0, 0, None, collections.Counter()])
for i, r in enumerate(data):
    k = int(r[ie]); g = c[k]; g[0] += 1; g[1] += int(r[sm]); g[2] += k
    if g[3] is None: g[3] = i
    for h in stalls: g[4][h[6:]] += int(r[hdr.index(h)])
tot = sum(v[2] for v in c.values()); ts = sum(v[1] for v in c.values())
print(f"total warp-instructions {tot}, samples {ts}, sass {len(data)}")
for k, v in sorted(c.items(), key=lambda x: -x[1][1])[:top]:
    why = " ".join(f"{n}:{x/max(v[1],1):.0%}" for n, x in v[4].most_common(4))
    print(f"exec {k:10d} n {v[0]:4d} samples {v[1]/ts:6.2%} instr {v[2]/tot:6.2%} idx {v[3]:5d}  {why}")
