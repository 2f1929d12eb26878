#!/bin/bash
# A/B the kernel variants under build/variants/ on config 2 (resident value, K1 ms, sizing-pass ms).
for so in pure_zlib_b200/libpzcuda.so build/variants/*.so; do
  PZ_LIBPZCUDA=$PWD/$so timeout 300 python bench.py --steps ${AB_STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e --verify 4 ${AB_ARGS} 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$so', round(d['value'],1), 'GB/s  K1', round(d['roofline']['kernel_ms'],2), 'ms  sizing', d['roofline']['decoder_only_ms'])"
done
