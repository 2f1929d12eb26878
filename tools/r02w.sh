#!/bin/bash
# round 2, GPU call W: K5 second version (branch-free symbol decode, tables in shared memory, input word loaded one refill ahead)
o=gpurun_out; tag=r02w
timeout 900 python -m pytest tests -m gpu -x -q -k "small_stream_batch or baseline_config or sizing or golden or mixed_verdicts" 2>&1 | tail -5 > $o/${tag}_pytest_k5.log
tail -2 $o/${tag}_pytest_k5.log
timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 64 > $o/${tag}_bench_records4k.json 2> $o/${tag}_bench_records4k.err
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pz_fixed_kernel -s 6 -c 3 --csv --log-file $o/${tag}_k5.csv python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_ncu.log 2>&1
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02w_bench_records4k.json").read().strip().splitlines()[-1])
print("value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"], b["checks"])
PY
grep "pz_fixed_kernel" $o/${tag}_k5.csv | awk -F'","' '{print $5, $(NF-2), $NF}'
