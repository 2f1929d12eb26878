#!/bin/bash
# GPU call: 2 / 3 symbols per trip on every config
o=gpurun_out; tag=r02ab
for v in trip2 trip3; do
  lib=pure_zlib_b200/libpzcuda_$v.so
  for cfg in text256k records4k huge text256k_l1; do
    PZ_LIBPZCUDA=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --config $cfg --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_${cfg}_$v.json 2> $o/${tag}_${cfg}_$v.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ab_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],3), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e)
PY
