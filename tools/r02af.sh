#!/bin/bash
# GPU call (2 GPUs): the torchrun path and the in-library multi-device path on the final tree
o=gpurun_out; tag=r02af
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $o/${tag}_bench_text256k_g2.json 2> $o/${tag}_g2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > $o/${tag}_bench_reference_g2.json 2> $o/${tag}_ref_g2.err
timeout 600 python -m pytest tests -m gpu -x -q -k "in_library_multi_gpu" 2>&1 | tail -2 > $o/${tag}_pytest_multi.log; tail -1 $o/${tag}_pytest_multi.log
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r02af_bench_text256k_g2.json").read().strip().splitlines()[-1])
print("g2 value", round(b["value"],1), b["scaling"], "n_gpus", b["n_gpus"], "e2e", b["e2e"]["value"], "ceiling", b["e2e"].get("ceiling"))
r=json.loads(open("gpurun_out/r02af_bench_reference_g2.json").read().strip().splitlines()[-1]); print("ref", r["value"], r.get("impl"), r["n_gpus"])
PY
