#!/bin/bash
# A/B the kernel variants under build/variants/ on config 5 (512 x 16 MiB stored): resident value and ms/step.
for so in pure_zlib_b200/libpzcuda.so build/variants/*.so; do
  PZ_LIBPZCUDA=$PWD/$so timeout 300 python bench.py --config stored16m --steps ${AB_STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e --verify 4 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$so', round(d['value'],1), 'GB/s ', round(d['ms_per_step'],3), 'ms/step  K2+K1', round(d['roofline']['kernel_ms'],3))"
done
