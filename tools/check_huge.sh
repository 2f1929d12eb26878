#!/bin/bash
# Config 4 after a K4 change, in one GPU call (tag r01k): the whole GPU suite, the bench line, the timeline, the launch list.
o=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee $o/r01k_pytest_gpu.log
timeout 600 python bench.py --config huge --steps 8 --warmup 3 > $o/r01k_bench_huge.json 2> /dev/null; cut -c1-160 $o/r01k_bench_huge.json
PZ_TRACE=1 timeout 400 python bench.py --config huge --steps 1 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep pz-k4 | tail -5 | tee $o/r01k_huge_timeline.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/r01k_launches_huge.csv python bench.py --config huge --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --verify 0 > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "huge_stream_block" 2>&1 | tail -4 | tee $o/r01k_memcheck_huge.log
