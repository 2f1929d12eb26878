#!/bin/bash
# round 2, GPU call A: the suite, the new bench line (all BASELINE workloads), the chain-only timing experiment
o=gpurun_out; tag=r02a
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $o/${tag}_box.txt; nproc >> $o/${tag}_box.txt; free -g | head -2 >> $o/${tag}_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $o/${tag}_pytest_gpu.log
( time timeout 1500 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err ) 2> $o/${tag}_bench.time
PZ_BENCH_NOCHECK=1 PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda_chain.so timeout 600 python bench.py --steps 10 --warmup 3 --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_bench_chain.json 2> $o/${tag}_bench_chain.err
tail -c 600 $o/${tag}_pytest_gpu.log; cut -c1-300 $o/${tag}_bench.json; cat $o/${tag}_bench.time
python - <<'PY'
import json
for f in ("gpurun_out/r02a_bench.json","gpurun_out/r02a_bench_chain.json"):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(b["value"],1), "k1", b["roofline"]["kernel_ms"], "dec", b["roofline"]["decoder_only_ms"], "e2e", b.get("e2e",{}).get("value"), "shim", b.get("e2e_shim",{}).get("value"))
        for k,v in b.get("other_configs",{}).items():
            print("  ",k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","skipped","failed","seconds")}, "e2e", v.get("e2e",{}).get("value"), "frac", v.get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
