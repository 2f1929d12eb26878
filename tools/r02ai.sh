#!/bin/bash
# GPU call: the writers' stores alone compiled out (their loads stay): which half of the copy costs the 1.1 ms
o=gpurun_out; tag=r02ai
PZ_BENCH_NOCHECK=1 PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda_nostore.so timeout 600 python bench.py --steps 10 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_bench_text256k_nostore.json 2> $o/${tag}_nostore.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02ai_bench_text256k_nostore.json").read().strip().splitlines()[-1])
print("nostore value", round(b["value"],1), "k1", round(b["roofline"]["kernel_ms"],3))
PY
