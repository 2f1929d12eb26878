#!/bin/bash
# round 2, GPU call K: K5 with the register bit buffer + K6 (dynamic blocks, tables in local memory) -- tests, records4k A/B
o=gpurun_out; tag=r02k
timeout 900 python -m pytest tests -m gpu -x -q -k "small_stream_batch or baseline_config or sizing or golden or mixed_verdicts" 2>&1 | tail -12 > $o/${tag}_pytest_k56.log
tail -4 $o/${tag}_pytest_k56.log
for b in 1 2 4; do
PZ_K6_BLOCKS=$b timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 64 > $o/${tag}_bench_records4k_b$b.json 2> $o/${tag}_bench_records4k_b$b.err
done
PZ_NO_K6=1 timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k_nok6.json 2> $o/${tag}_bench_records4k_nok6.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file $o/${tag}_launches_records4k.csv python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_launches.log 2>&1
python - <<'PY'
import json
for v in ("_b1","_b2","_b4","_nok6"):
    try:
        b=json.loads(open(f"gpurun_out/r02k_bench_records4k{v}.json").read().strip().splitlines()[-1])
        print(v, "value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(v, "ERR", e, open(f"gpurun_out/r02k_bench_records4k{v}.err").read()[-400:])
PY
grep -E "pz_" $o/${tag}_launches_records4k.csv | tail -9 | awk -F'","' '{print $5, $NF}'
