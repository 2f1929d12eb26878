#!/bin/bash
o=gpurun_out; tag=r02l
for cfg in "1 256" "1 128" "1 64"; do set -- $cfg
PZ_K6_BLOCKS=$1 PZ_K6_THREADS=$2 timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k_b$1_t$2.json 2> $o/${tag}_bench_records4k_b$1_t$2.err
done
PZ_K6_BLOCKS=1 PZ_K6_THREADS=128 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file $o/${tag}_launches_records4k.csv python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_launches.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02l_bench_records4k_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[-16:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e)
PY
grep -E "pz_" $o/${tag}_launches_records4k.csv | tail -9 | awk -F'","' '{print $5, $NF}'
