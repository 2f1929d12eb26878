#!/bin/bash
# Final check of the round in one GPU call (tag r01j): whole GPU suite, smoke, config 2 and config 4 lines, memcheck of the
# paths that changed after r01h, launch list of config 4.
o=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $o/r01j_pytest_gpu.log; tail -2 $o/r01j_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > $o/r01j_bench_text256k.json 2> /dev/null; cut -c1-160 $o/r01j_bench_text256k.json
timeout 600 python bench.py --config huge --steps 8 --warmup 3 > $o/r01j_bench_huge.json 2> /dev/null; cut -c1-160 $o/r01j_bench_huge.json
PZ_TRACE=1 timeout 400 python bench.py --config huge --steps 1 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep pz-k4 | tail -5 | tee $o/r01j_huge_timeline.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/r01j_launches_huge.csv python bench.py --config huge --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --verify 0 > /dev/null 2>&1
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "feed or incremental or pump or multichunk or huge_stream_block" 2>&1 | tail -6 > $o/r01j_memcheck.log; cat $o/r01j_memcheck.log
