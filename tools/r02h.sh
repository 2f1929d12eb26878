#!/bin/bash
# round 2, GPU call H: one vote and one branch per trip in the hot warp -- suite, then A/B against the loop before
o=gpurun_out; tag=r02h
timeout 180 python -m pytest tests -m gpu -x -q -k "golden or appendix_b or kat" 2>&1 | tail -3 > $o/${tag}_smoke.log
if ! grep -q "passed" $o/${tag}_smoke.log || grep -q "failed" $o/${tag}_smoke.log; then cat $o/${tag}_smoke.log; echo SMOKE FAILED; exit 1; fi
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $o/${tag}_pytest_gpu.log ) 2> $o/${tag}_pytest.time
tail -3 $o/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
PZ_NO_LEAN=1 PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda_oldloop.so timeout 300 python bench.py --steps 20 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_oldloop.json 2> $o/${tag}_bench_oldloop.err
timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k.json 2> $o/${tag}_bench_records4k.err
timeout 600 python bench.py --steps 5 --warmup 3 --config huge --others none --no-e2e --no-cpu-baseline > $o/${tag}_bench_huge.json 2> $o/${tag}_bench_huge.err
python - <<'PY'
import json
for v in ("","_oldloop","_records4k","_huge"):
    try:
        b=json.loads(open(f"gpurun_out/r02h_bench{v}.json").read().strip().splitlines()[-1])
        print(v or "newloop", "value", round(b["value"],1), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(v, "ERR", e, open(f"gpurun_out/r02h_bench{v}.err").read()[-300:])
PY
