#!/bin/bash
# One GPU call: the whole GPU suite, then the K4 timeline.
o=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $o/r01h_pytest_gpu.log
tail -4 $o/r01h_pytest_gpu.log
PZ_TRACE=1 timeout 400 python bench.py --config huge --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $o/r01h_huge_onepass.json 2> $o/r01h_huge_onepass.err
grep "pz-k4" $o/r01h_huge_onepass.err | tail -6; cut -c1-260 $o/r01h_huge_onepass.json
