#!/bin/bash
# GPU call: merged room check (adv <= room); A/B of the window-base update written as arithmetic
o=gpurun_out; tag=r02ae
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or appendix or mixed_verdicts or baseline_config or large_expansion or output" 2>&1 | tail -3 > $o/${tag}_pytest.log; tail -1 $o/${tag}_pytest.log
for v in base; do
  lib=pure_zlib_b200/libpzcuda_$v.so; [ $v = base ] && lib=pure_zlib_b200/libpzcuda.so
  PZ_LIBPZCUDA=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_text256k_$v.json 2> $o/${tag}_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ae_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],3), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e)
PY
