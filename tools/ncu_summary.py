#!/usr/bin/env python
"""Reduce an ncu report (--set full) to a small JSON summary: one object per captured launch with
the metrics the roofline discussion uses.  usage: ncu_summary.py report.ncu-rep out.json [traffic.json config [kernel-substring]]
With the last two arguments the DRAM traffic of the first pz_inflate_kernel launch is also
written into profiles/roofline_traffic.json under `config` (read by bench.py)."""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
res = []
for r in data:
    o = {"Kernel Name": r[hdr.index("Kernel Name")], "Grid Size": r[hdr.index("Grid Size")], "Block Size": r[hdr.index("Block Size")]}
    u = {}
    for k in KEYS:
        if k in hdr:
            o[k] = r[hdr.index(k)]; u[k] = units[hdr.index(k)]
    o["units"] = u
    res.append(o)
json.dump(res, open(out, "w"), indent=1)
print(f"{len(res)} launches -> {out}")
if len(sys.argv) > 4:
    tpath, cfg = sys.argv[3], sys.argv[4]
    try:
        t = json.load(open(tpath))
    except Exception:
        t = {}
    want = sys.argv[5] if len(sys.argv) > 5 else "pz_inflate_kernel"
    k1 = next(o for o in res if want in o["Kernel Name"])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = float(k1["dram__bytes_read.sum"]) * scale[k1["units"]["dram__bytes_read.sum"]]
    wr = float(k1["dram__bytes_write.sum"]) * scale[k1["units"]["dram__bytes_write.sum"]]
    t[cfg] = {"kernel": k1["Kernel Name"], "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
              "source": f"{out} (ncu --set full, one launch)"}
    json.dump(t, open(tpath, "w"), indent=1)
