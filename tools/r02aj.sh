#!/bin/bash
# GPU call (8 GPUs): the scaling line at N = 8 with the final kernels
o=gpurun_out; tag=r02aj
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > $o/${tag}_bench_text256k_g8.json 2> $o/${tag}_g8.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02aj_bench_text256k_g8.json").read().strip().splitlines()[-1])
print("g8 value", round(b["value"],1), b["scaling"], "ms", round(b["ms_per_step"],3), "e2e", round(b["e2e"]["value"],1), "ceiling", b["e2e"].get("ceiling"))
PY
