#!/bin/bash
# GPU call: LUT entry fields as ready-made shift amounts (no mask instructions between the tables and the shifts)
o=gpurun_out; tag=r02ag
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or appendix or mixed_verdicts or baseline_config or large_expansion or huge_stream_block or oracle_mixed" 2>&1 | tail -3 > $o/${tag}_pytest.log; tail -1 $o/${tag}_pytest.log
for cfg in text256k records4k huge text256k_l1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --config $cfg --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_$cfg.json 2> $o/${tag}_$cfg.err
done
PZ_LEAN=1 timeout 600 python bench.py --steps 5 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 64 > $o/${tag}_bench_text256k_lean.json 2> $o/${tag}_lean.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ag_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[23:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],3), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"], b["checks"])
    except Exception as e: print(f, "ERR", e)
PY
