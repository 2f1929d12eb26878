#!/bin/bash
o=gpurun_out; tag=r02d
timeout 300 python tools/debug_gzip.py > $o/${tag}_debug_gzip.txt 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q -k "gzip or crc32 or memory_is_bounded" 2>&1 | tail -30 > $o/${tag}_pytest_new.log ) 2> $o/${tag}_pytest.time
for v in "" _mad; do
  PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda$v.so timeout 600 python bench.py --steps 20 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 8 > $o/${tag}_bench$v.json 2> $o/${tag}_bench$v.err
done
cat $o/${tag}_debug_gzip.txt | tail -30; tail -12 $o/${tag}_pytest_new.log; cat $o/${tag}_pytest.time
python - <<'PY'
import json
for v in ("","_mad"):
    try:
        b=json.loads(open(f"gpurun_out/r02d_bench{v}.json").read().strip().splitlines()[-1])
        print(v or "base", "value", round(b["value"],1), "k1", round(b["roofline"]["kernel_ms"],3), "dec", round(b["roofline"]["decoder_only_ms"],3))
    except Exception as e: print(v, "ERR", e)
PY
