#!/bin/bash
# round 2, GPU call Y: the deep fuzz, six rounds of 250 000 cases with fresh seeds
o=gpurun_out; tag=r02y
: > $o/${tag}_fuzz.jsonl
for r in 1 2 3 4 5 6; do
  timeout 900 python tools/fuzz_gpu.py --seeds 100 --per-seed 2500 --first-seed $((20000 + r * 1000)) --incremental 5000 >> $o/${tag}_fuzz.jsonl 2>> $o/${tag}_fuzz.err; echo "round $r rc=$?"
done
python - <<'PY'
import json
t=[json.loads(l) for l in open("gpurun_out/r02y_fuzz.jsonl")]
print("cases", sum(x["cases"] for x in t), "compared", sum(x["compared"] for x in t), "mismatches", sum(x["mismatches"] for x in t), "incremental", sum(x["incremental_cases"] for x in t), "inc mismatches", sum(x["incremental_mismatches"] for x in t), "quirk", sum(x["incremental_skipped_chunk_boundary_quirk"] for x in t))
for x in t:
    for m in x["first_mismatches"][:3]: print(json.dumps(m)[:600])
PY
# big streams: 40 seeds x 500 mutations of 10 KB .. 400 KB streams (20 000 cases, ~2 GB decoded per entry point)
timeout 1200 python tools/fuzz_gpu.py --big --seeds 40 --per-seed 500 --first-seed 300 --incremental 1500 > $o/${tag}_fuzz_big.json 2>> $o/${tag}_fuzz.err; echo "big rc=$?"
cut -c1-700 $o/${tag}_fuzz_big.json
