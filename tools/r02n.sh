#!/bin/bash
# round 2, GPU call N: K5 / K6 with packed stream lists and the merged decode + produce loop
o=gpurun_out; tag=r02n
timeout 900 python -m pytest tests -m gpu -x -q -k "small_stream_batch or baseline_config or sizing or golden or mixed_verdicts" 2>&1 | tail -12 > $o/${tag}_pytest_k56.log
tail -3 $o/${tag}_pytest_k56.log
for cfg in "8 4" "4 4" "2 4" "8 2"; do set -- $cfg
PZ_K5_BLOCKS=$1 PZ_K6_BLOCKS=$2 timeout 600 python bench.py --steps 5 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 16 > $o/${tag}_bench_records4k_k5b$1_k6b$2.json 2> $o/${tag}_bench_records4k_k5b$1_k6b$2.err
done
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file $o/${tag}_launches_records4k.csv python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_launches.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02n_bench_records4k_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[-18:], "value", round(b["value"],1), "ms", round(b["ms_per_step"],2), "k1", round(b["roofline"]["kernel_ms"],3), "dec", b["roofline"]["decoder_only_ms"])
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
grep -E "pz_" $o/${tag}_launches_records4k.csv | tail -44 | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -44
