#!/bin/bash
# GPU call: symbols per trip of the hot warp after the round-2 loop head (one vote, one branch per trip): 3 / 4 / 5 / 6
o=gpurun_out; tag=r02aa
for v in base trip3 trip5 trip6; do
  lib=pure_zlib_b200/libpzcuda_$v.so; [ $v = base ] && lib=pure_zlib_b200/libpzcuda.so
  PZ_LIBPZCUDA=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_text256k_$v.json 2> $o/${tag}_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02aa_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[22:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],3), "k1", round(b["roofline"]["kernel_ms"],3), "dec", round(b["roofline"]["decoder_only_ms"],3))
    except Exception as e: print(f, "ERR", e)
PY
