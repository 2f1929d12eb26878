#!/bin/bash
# Everything profiles/ holds for one round, in one GPU call.  usage: tools/profile_round.sh r01f
# (run under gpurun from the repo root; results land in gpurun_out/ and are copied to profiles/ by hand)
tag=${1:-rXX}
o=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_text256k.json 2> $o/${tag}_bench_text256k.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_text256k.json 2> $o/${tag}_bench_reference.err
for c in records4k stored16m huge; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > $o/${tag}_bench_$c.json 2> $o/${tag}_bench_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_text256k.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_launches.log 2>&1
# one step of config 2 (K2 probe/copy, K1, K3a, K3b), then the K1-only launch and the sizing pass
ncu --set full --clock-control none --import-source on -k regex:pz_ -s 15 -c 8 -o $o/${tag}_k1_k3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --verify 0 > $o/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pz_stored_copy -s 4 -c 1 -o $o/${tag}_k2 \
  python bench.py --config stored16m --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --verify 0 > $o/${tag}_ncu_k2.log 2>&1
tail -1 $o/${tag}_bench_text256k.json | cut -c1-400
