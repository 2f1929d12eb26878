#!/bin/bash
# One GPU call: the whole GPU suite, the incremental bench, the K4 timeline (one pass against two).
o=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $o/r01h_pytest_gpu.log
tail -6 $o/r01h_pytest_gpu.log
timeout 300 python tools/bench_incremental.py --streams 1024 --pieces 8 > $o/r01h_bench_incremental.json 2> $o/r01h_bench_incremental.err
tail -3 $o/r01h_bench_incremental.err; cat $o/r01h_bench_incremental.json
timeout 300 python tools/bench_incremental.py --streams 4096 --pieces 32 > $o/r01h_bench_incremental_4096x32.json 2> $o/r01h_bench_incremental2.err
tail -3 $o/r01h_bench_incremental2.err; cat $o/r01h_bench_incremental_4096x32.json
PZ_TRACE=1 timeout 400 python bench.py --config huge --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $o/r01h_huge_onepass.json 2> $o/r01h_huge_onepass.err
grep "pz-k4" $o/r01h_huge_onepass.err | tail -6; cut -c1-260 $o/r01h_huge_onepass.json
PZ_K4_TWO_PASS=1 timeout 400 python bench.py --config huge --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $o/r01h_huge_twopass.json 2> $o/r01h_huge_twopass.err
cut -c1-260 $o/r01h_huge_twopass.json
