#!/bin/bash
# round 2, GPU call R: K5 with few threads per SM (does the history of the streams in flight fit the L2?)
o=gpurun_out; tag=r02r
for cfg in "1 64" "1 128" "1 192" "1 256" "2 256" "8 256"; do set -- $cfg
PZ_NO_K6=1 PZ_K5_BLOCKS=$1 PZ_K5_THREADS=$2 timeout 600 python bench.py --steps 3 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k_b$1_t$2.json 2> $o/${tag}_bench_records4k_b$1_t$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02r_bench_records4k_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[32:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
