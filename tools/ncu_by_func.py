#!/usr/bin/env python
"""Attribute the samples of an ncu report of pz_inflate_kernel<false> to source functions of
pz_device.cuh (via nvdisasm line info of the cubin inside libpzcuda.so; the .so must be the
build that was profiled).  usage: ncu_by_func.py report.ncu-rep"""
import csv, subprocess, io, sys, re, os, tempfile, collections, bisect
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "pure_zlib_b200", "libpzcuda.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = os.path.join(tmp, "pz_kernels.sm_100a.cubin")
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cub], capture_output=True, text=True).stdout
if not dis:
    dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
# function line ranges in pz_device.cuh
dev = os.path.join(root, "pure_zlib_b200", "csrc", "pz_device.cuh")
starts = []
for n, line in enumerate(open(dev), 1):
    m = re.match(r"(?:PZ_DEV|PZ_COLD).*?\b(pz_\w+)\s*\(", line)
    if m:
        starts.append((n, m.group(1)))
def func_of(path, line):
    if not path.endswith("pz_device.cuh"): return os.path.basename(path)
    i = bisect.bisect_right([s for s, _ in starts], line) - 1
    return starts[i][1] if i >= 0 else "?"
# walk the disassembly of the <false> kernel
lines = dis.splitlines()
kern = "ILb1ELb0ELb0E" if len(sys.argv) > 2 and sys.argv[2] == "count" else "ILb0ELb0ELb0E"
i0 = next(i for i, l in enumerate(lines) if l.startswith(".text._Z17pz_inflate_kernel" + kern))
ONLY_KERNEL = False
chain = []; last = [("?", 0)]; per_instr = []
for l in lines[i0 + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        chain.append((m.group(1), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
        if chain: last = chain
        per_instr.append(last)
        chain = []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]
nxt = next((i for i, r in enumerate(rows[2:], 2) if r and r[0] == "Kernel Name"), len(rows))  # a report with several launches: the first one
data = [r for r in rows[2:nxt] if len(r) == len(hdr)]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
print(f"sass instrs: ncu {len(data)}, nvdisasm {len(per_instr)}")
TOP = {"pz_decoder_warp", "pz_writer_warp", "pz_slow_step", "pz_fast_loop"}
for title, pick in (("innermost function", lambda fs: fs[0]),
                    ("phase (outermost function below the warp loops)", lambda fs: next((f for f in reversed(fs) if f not in TOP and f.startswith("pz_")), fs[-1]))):
    agg = collections.defaultdict(lambda: [0, 0])
    for r, ch in zip(data, per_instr):
        fs = [func_of(p_, l_) for p_, l_ in ch]
        f = pick(fs)
        agg[f][0] += int(r[sm]); agg[f][1] += int(r[ie])
    ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
    print("--", title)
    for f, (s_, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
        print(f"{f:28s} samples {s_/ts:6.2%}  instr {n/ti:6.2%}")
