#!/bin/bash
# round 2, GPU call G: where does the lean kernel wait?  ncu source counters of one launch + a no-copy timing variant
o=gpurun_out; tag=r02g
PZ_BENCH_NOCHECK=1 PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda_nocopy.so timeout 300 python bench.py --steps 10 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_bench_nocopy.json 2> $o/${tag}_bench_nocopy.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pz_inflate_kernel -s 8 -c 2 -o $o/${tag}_k1lean python bench.py --steps 1 --warmup 3 --others none --no-cpu-baseline --no-e2e --verify 0 > $o/${tag}_ncu.log 2>&1
ls -la $o/${tag}_k1lean.ncu-rep
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02g_bench_nocopy.json").read().strip().splitlines()[-1])
print("nocopy lean", "k1", round(b["roofline"]["kernel_ms"],3))
PY
