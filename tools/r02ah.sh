#!/bin/bash
# GPU call: what separates the decode kernel (6.65 ms) from the sizing pass (5.27 ms) now: writers without their copies; writer nap length
o=gpurun_out; tag=r02ah
for v in nocopy nap250 nap4; do
  PZ_BENCH_NOCHECK=1 PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_bench_text256k_$v.json 2> $o/${tag}_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ah_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],3), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e)
PY
