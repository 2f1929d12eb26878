#!/bin/bash
# round 2, GPU call I (8 GPUs): host-side copy ceiling, one-process multi-GPU, the scaling line at N = 8 (weak) and records4k sharded + gathered
o=gpurun_out; tag=r02i
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > $o/${tag}_box.txt; nproc >> $o/${tag}_box.txt
timeout 600 python tools/pcie_ceiling.py --gpus 1,2,4,8 > $o/${tag}_pcie_ceiling.jsonl 2> $o/${tag}_pcie.err
timeout 900 python tools/bench_inlib_multigpu.py --config text256k --gpus 1,2,4,8 > $o/${tag}_inlib_text256k.jsonl 2> $o/${tag}_inlib.err
timeout 900 python tools/bench_inlib_multigpu.py --config records4k --gpus 1,2,4,8 > $o/${tag}_inlib_records4k.jsonl 2>> $o/${tag}_inlib.err
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > $o/${tag}_bench_text256k_g$n.json 2> $o/${tag}_g$n.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --config records4k --shard --gather > $o/${tag}_bench_records4k_shard$n.json 2> $o/${tag}_shard$n.err
done
cat $o/${tag}_pcie_ceiling.jsonl $o/${tag}_inlib_text256k.jsonl $o/${tag}_inlib_records4k.jsonl | cut -c1-260
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i_bench_*.json")):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(b["value"],1), b["scaling"], "e2e", round(b.get("e2e",{}).get("value",0),1), "gather", (b.get("gather") or {}).get("seconds"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 2 $o/${tag}_inlib.err $o/${tag}_shard8.err
