#!/bin/bash
# round 2, GPU call J: K5 (one thread per small fixed-Huffman stream) -- tests, then records4k with and without it
o=gpurun_out; tag=r02j
timeout 900 python -m pytest tests -m gpu -x -q -k "small_stream_batch or baseline_config or sizing or golden" 2>&1 | tail -12 > $o/${tag}_pytest_k5.log
tail -4 $o/${tag}_pytest_k5.log
timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-cpu-baseline --verify 64 > $o/${tag}_bench_records4k.json 2> $o/${tag}_bench_records4k.err
PZ_NO_K5=1 timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k_nok5.json 2> $o/${tag}_bench_records4k_nok5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/${tag}_launches_records4k.csv python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_launches.log 2>&1
python - <<'PY'
import json
for v in ("","_nok5"):
    try:
        b=json.loads(open(f"gpurun_out/r02j_bench_records4k{v}.json").read().strip().splitlines()[-1])
        print(v or "k5", "value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"], "e2e", b.get("e2e",{}).get("value"))
    except Exception as e: print(v, "ERR", e, open(f"gpurun_out/r02j_bench_records4k{v}.err").read()[-400:])
PY
grep -E "pz_" $o/${tag}_launches_records4k.csv | tail -8 | cut -c1-220
