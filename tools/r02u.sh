#!/bin/bash
# round 2, GPU call U: the pump after batching the history moves (one kernel), O(n) pump list, later input give-back
o=gpurun_out; tag=r02u
timeout 900 python -m pytest tests -m gpu -x -q -k "incremental or pump or feed or multichunk or deflate_cli or concurrent or gzip" 2>&1 | tail -5 > $o/${tag}_pytest_incremental.log
tail -2 $o/${tag}_pytest_incremental.log
timeout 600 python tools/bench_incremental.py --streams 4096 --pieces 32 > $o/${tag}_bench_incremental_4096x32.json 2> $o/${tag}_bench_incremental.err
timeout 600 python tools/bench_incremental.py --streams 1024 --pieces 8 > $o/${tag}_bench_incremental_1024x8.json 2>> $o/${tag}_bench_incremental.err
python - <<'PY'
import json
for f in ("gpurun_out/r02u_bench_incremental_4096x32.json","gpurun_out/r02u_bench_incremental_1024x8.json"):
    b=json.load(open(f))
    for k in ("resumed","resumed_single_feeds"):
        r=b[k]; print(f[-12:],k, round(r["value"],2),"GB/s pump total",r["pump_ms_total"],"feed",r["feed_ms"],"drain",r["drain_ms"], r["pump_ms"][:12], r["memory"])
PY
