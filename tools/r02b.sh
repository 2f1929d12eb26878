#!/bin/bash
# round 2, GPU call B (2 GPUs): in-library multi-GPU, NCCL gather of sharded outputs, PCIe ceiling
o=gpurun_out; tag=r02b
nvidia-smi --query-gpu=name --format=csv,noheader > $o/${tag}_box.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "decompress_batch_single or multi_gpu or concurrent or golden" 2>&1 | tail -8 > $o/${tag}_pytest.log
timeout 600 python tools/pcie_ceiling.py --gpus 1,2 > $o/${tag}_pcie_ceiling.jsonl 2> $o/${tag}_pcie.err
timeout 900 python tools/bench_inlib_multigpu.py --config text256k --gpus 1,2 > $o/${tag}_inlib_text256k.jsonl 2> $o/${tag}_inlib.err
timeout 900 python tools/bench_inlib_multigpu.py --config records4k --gpus 1,2 > $o/${tag}_inlib_records4k.jsonl 2>> $o/${tag}_inlib.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --config records4k --shard --gather > $o/${tag}_bench_records4k_shard2.json 2> $o/${tag}_shard.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > $o/${tag}_bench_text256k_g2.json 2> $o/${tag}_g2.err
cat $o/${tag}_pytest.log | tail -4; cat $o/${tag}_pcie_ceiling.jsonl $o/${tag}_inlib_text256k.jsonl $o/${tag}_inlib_records4k.jsonl | cut -c1-400
python - <<'PY'
import json
for f in ("gpurun_out/r02b_bench_records4k_shard2.json","gpurun_out/r02b_bench_text256k_g2.json"):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(b["value"],1), b["scaling"], "e2e", b.get("e2e",{}).get("value"), "gather", b.get("gather"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $o/${tag}_shard.err $o/${tag}_inlib.err
