#!/bin/bash
o=gpurun_out; tag=r02m
PZ_K6_BLOCKS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pz_fixed_kernel -s 4 -c 4 -o $o/${tag}_k56 python bench.py --steps 1 --warmup 3 --config records4k --others none --no-cpu-baseline --no-e2e --verify 0 > $o/${tag}_ncu.log 2>&1
ls -la $o/${tag}_k56.ncu-rep
