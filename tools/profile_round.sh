#!/bin/bash
# Everything profiles/ holds for one round, in one GPU call.  usage: tools/profile_round.sh r01h
# (run under gpurun from the repo root; results land in gpurun_out/ and are copied to profiles/ by hand)
tag=${1:-rXX}
o=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $o/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_text256k.json 2> $o/${tag}_bench_text256k.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_text256k.json 2> $o/${tag}_bench_reference.err
for c in records4k stored16m huge; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > $o/${tag}_bench_$c.json 2> $o/${tag}_bench_$c.err
done
PZ_K4_TWO_PASS=1 timeout 600 python bench.py --config huge --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $o/${tag}_bench_huge_twopass.json 2> /dev/null
PZ_TRACE=1 timeout 600 python bench.py --config huge --steps 1 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 > /dev/null | grep "pz-k4" | tail -5 > $o/${tag}_huge_timeline.txt
timeout 300 python tools/bench_incremental.py --streams 1024 --pieces 8 > $o/${tag}_bench_incremental_1024x8.json 2> $o/${tag}_bench_incremental.err
timeout 300 python tools/bench_incremental.py --streams 4096 --pieces 32 > $o/${tag}_bench_incremental_4096x32.json 2>> $o/${tag}_bench_incremental.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_text256k.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_launches.log 2>&1
# one step of config 2 (K2 probe/copy, K1, K3a, K3b), then the K1-only launch and the sizing pass
ncu --set full --clock-control none --import-source on -k regex:pz_ -s 15 -c 8 -o $o/${tag}_k1_k3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --verify 0 > $o/${tag}_ncu.log 2>&1
python tools/ncu_summary.py $o/${tag}_k1_k3.ncu-rep $o/${tag}_ncu_k1_k3_summary.json > $o/${tag}_ncu_summary.log 2>&1
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "incremental or pump or multichunk or huge_stream_block" 2>&1 | tail -12 > $o/${tag}_memcheck.log
tail -1 $o/${tag}_bench_text256k.json | cut -c1-400
