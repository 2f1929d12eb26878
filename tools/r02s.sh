#!/bin/bash
# round 2, GPU call S: validation of the tree as it stands -- full GPU suite, smoke, the default bench line (all configs), the reference arm,
# launch list of the default command, compute-sanitizer over the parity subset that covers the round's new code paths
o=gpurun_out; tag=r02s
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $o/${tag}_pytest_gpu.log
tail -3 $o/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout 1500 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err
timeout 900 python bench.py --impl reference > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02s_bench_default.json").read().strip().splitlines()[-1])
print("headline", round(b["value"],1), "ms", round(b["ms_per_step"],3), "frac", round(b["roofline"]["frac"],4), "e2e", b["e2e"], "launches", b["gpu_launches"])
print("cpu_baseline", b["cpu_baseline"])
for k,v in b.get("other_configs",{}).items():
    print(k, {kk: (round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","skipped","note")}, "e2e", (v.get("e2e") or {}).get("value"), "frac", (v.get("roofline") or {}).get("frac"))
for k in ("e2e_shim","haskell"): print(k, b.get(k))
r=json.loads(open("gpurun_out/r02s_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["unit"], r["cpu_baseline"])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "gzip or incremental_matches or mixed_verdicts or golden or appendix" > $o/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $o/${tag}_memcheck.log
