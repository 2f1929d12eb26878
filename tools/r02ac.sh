#!/bin/bash
# GPU call: validation of the tree with PZ_TRIP = 3 (full suite, smoke, fuzz, default bench, reference arm, memcheck subset, ncu summary of one step of config 2)
o=gpurun_out; tag=r02ac
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $o/${tag}_pytest_gpu.log
tail -3 $o/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout 900 python tools/fuzz_gpu.py --seeds 100 --per-seed 2500 --first-seed 40000 --incremental 5000 > $o/${tag}_fuzz.json 2> $o/${tag}_fuzz.err; echo "fuzz rc=$?"
timeout 900 python tools/fuzz_gpu.py --big --seeds 24 --per-seed 300 --first-seed 900 --incremental 1000 > $o/${tag}_fuzz_big.json 2>> $o/${tag}_fuzz.err; echo "fuzz big rc=$?"
timeout 900 python tools/fuzz_gpu.py --big --huge-bytes 16384 --seeds 12 --per-seed 300 --first-seed 950 --incremental 0 > $o/${tag}_fuzz_k4.json 2>> $o/${tag}_fuzz.err; echo "fuzz k4 rc=$?"
timeout 1500 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err
timeout 900 python bench.py --impl reference > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02ac_bench_default.json").read().strip().splitlines()[-1])
print("headline", round(b["value"],1), "ms", round(b["ms_per_step"],3), "k1", round(b["roofline"]["kernel_ms"],3), "dec", round(b["roofline"]["decoder_only_ms"],3), "frac", round(b["roofline"]["frac"],4), "e2e", round(b["e2e"]["value"],1), "launches", b["gpu_launches"], "clocks", b.get("clocks"))
for k,v in b.get("other_configs",{}).items():
    print(k, {kk: (round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","skipped","note")}, "e2e", (v.get("e2e") or {}).get("value"), "frac", (v.get("roofline") or {}).get("frac"))
print("e2e_shim", b.get("e2e_shim",{}).get("value"))
for f in ("fuzz","fuzz_big","fuzz_k4"):
    x=json.loads(open(f"gpurun_out/r02ac_{f}.json").read()); print(f, x["cases"], "mismatches", x["mismatches"], "incremental", x["incremental_cases"], x["incremental_mismatches"], "k4", x.get("k4_done"), x.get("k4_declined"))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "gzip or incremental_matches or mixed_verdicts or golden or appendix or huge_stream_block" > $o/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $o/${tag}_memcheck.log
ncu --set full --clock-control none --import-source on -k regex:pz_ -s 15 -c 8 -f -o $o/${tag}_k1_k3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --verify 0 --others none > $o/${tag}_ncu_text256k.log 2>&1
python tools/ncu_summary.py $o/${tag}_k1_k3.ncu-rep $o/${tag}_ncu_k1_k3_summary.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/${tag}_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --others records4k,huge,stored16m > $o/${tag}_launches.log 2>&1
