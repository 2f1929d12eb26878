#!/bin/bash
# GPU call: e2e of config 2 with and without the one-launch sizing pass, same box (the host side differs from box to box)
o=gpurun_out; tag=r02al
for v in prev base prev base; do
  lib=pure_zlib_b200/libpzcuda_$v.so; [ $v = base ] && lib=pure_zlib_b200/libpzcuda.so
  PZ_LIBPZCUDA=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --others none --verify 0 2> /dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'value', round(b['value'],1), 'e2e', round(b['e2e']['value'],1), 'shim', round(b['e2e_shim']['value'],1))"
done
