#!/bin/bash
# GPU call: the sizing pass of a host-buffer batch in ONE launch fed in pieces (it was cut into eight slices): parity subset, the shim call, e2e
o=gpurun_out; tag=r02ak
timeout 900 python -m pytest tests -m gpu -x -q -k "sizing or decompress_batch or golden or mixed_verdicts or baseline_config or concurrent or gzip_and_raw_batch or stored" 2>&1 | tail -3 > $o/${tag}_pytest.log; tail -1 $o/${tag}_pytest.log
PZ_TRACE=1 timeout 300 python tools/trace_shim.py 2>&1 | grep -v issued | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --others none --verify 16 > $o/${tag}_bench_text256k.json 2> $o/${tag}_bench.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02ak_bench_text256k.json").read().strip().splitlines()[-1])
print("value", round(b["value"],1), "e2e", round(b["e2e"]["value"],1), "e2e_shim", b.get("e2e_shim",{}).get("value"))
PY
