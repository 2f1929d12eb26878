timeout 400 python -m pytest tests -m gpu -q -k "huge" 2>&1 | tail -2
timeout 400 python bench.py --config huge --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],2), 'GB/s', round(d['ms_per_step'],2), 'ms')"
PZ_TRACE=1 timeout 400 python bench.py --config huge --steps 1 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep pz-k4 | tail -5
