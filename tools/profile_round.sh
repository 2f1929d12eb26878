#!/bin/bash
# Everything profiles/ holds for one round, in one GPU call.  usage: tools/profile_round.sh r01b
# (run under gpurun from the repo root; results land in gpurun_out/ and are copied to profiles/ by hand)
tag=${1:-rXX}
o=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_text256k.json 2> $o/${tag}_bench_text256k.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_text256k.json 2> $o/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_text256k.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pz_ -s 15 -c 5 -o $o/${tag}_k1_k3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --verify 0 > $o/${tag}_ncu.log 2>&1
tail -1 $o/${tag}_bench_text256k.json | cut -c1-400
