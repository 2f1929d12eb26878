#!/bin/bash
# round 2, GPU call P: where do K1's service groups spend a small dynamic stream?  (ncu --set full with source, records4k without K6)
o=gpurun_out; tag=r02p
PZ_NO_K6=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:pz_inflate_kernel -c 4 -f -o $o/${tag}_k1_records4k python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_ncu.log 2>&1
tail -3 $o/${tag}_ncu.log
ls -la $o/${tag}_k1_records4k.ncu-rep
