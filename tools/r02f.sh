#!/bin/bash
# round 2, GPU call F: the lean kernel -- smoke first (a hang must not eat the budget), then the suite, then A/B
o=gpurun_out; tag=r02f
timeout 180 python -m pytest tests -m gpu -x -q -k "golden or appendix_b or kat" 2>&1 | tail -5 > $o/${tag}_smoke.log
if ! grep -q "passed" $o/${tag}_smoke.log || grep -q "failed" $o/${tag}_smoke.log; then cat $o/${tag}_smoke.log; echo SMOKE FAILED; exit 1; fi
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $o/${tag}_pytest_gpu.log ) 2> $o/${tag}_pytest.time
tail -4 $o/${tag}_pytest_gpu.log
for v in "" _trip4 _trip8; do
  PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda$v.so timeout 300 python bench.py --steps 20 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench$v.json 2> $o/${tag}_bench$v.err
done
PZ_NO_LEAN=1 timeout 300 python bench.py --steps 20 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_nolean.json 2> $o/${tag}_bench_nolean.err
timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k.json 2> $o/${tag}_bench_records4k.err
python - <<'PY'
import json
for v in ("","_trip4","_trip8","_nolean","_records4k"):
    try:
        b=json.loads(open(f"gpurun_out/r02f_bench{v}.json").read().strip().splitlines()[-1])
        print(v or "lean6", "value", round(b["value"],1), "k1", round(b["roofline"]["kernel_ms"],3), "dec", round(b["roofline"]["decoder_only_ms"],3))
    except Exception as e: print(v, "ERR", e, open(f"gpurun_out/r02f_bench{v}.err").read()[-300:])
PY
