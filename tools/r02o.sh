#!/bin/bash
# round 2, GPU call O: service groups of 16 lanes (two slots per service warp instead of four) -- records4k without K6, text256k
o=gpurun_out; tag=r02o
for v in base g16p3 g16w8; do
  lib=pure_zlib_b200/libpzcuda_$v.so; [ $v = base ] && lib=pure_zlib_b200/libpzcuda.so
  for cfg in records4k text256k; do
    PZ_NO_K6=1 PZ_LIBPZCUDA=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --config $cfg --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_${cfg}_$v.json 2> $o/${tag}_bench_${cfg}_$v.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02o_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
