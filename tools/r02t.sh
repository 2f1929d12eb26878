#!/bin/bash
# round 2, GPU call T: the profile set of the round's final kernels (launch lists, ncu --set full of K1 / K3 on config 2 and of K5 / K1 on config 3), incremental bench
o=gpurun_out; tag=r02t
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/${tag}_launches_default.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --others records4k,huge,stored16m > $o/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pz_ -s 15 -c 8 -f -o $o/${tag}_k1_k3 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --verify 0 --others none > $o/${tag}_ncu_text256k.log 2>&1
python tools/ncu_summary.py $o/${tag}_k1_k3.ncu-rep $o/${tag}_ncu_k1_k3_summary.json > $o/${tag}_ncu_summary.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pz_fixed_kernel|pz_inflate_kernel" -s 6 -c 2 -f -o $o/${tag}_k5_k1_records4k \
  python bench.py --steps 1 --warmup 3 --config records4k --no-cpu-baseline --no-e2e --verify 0 --others none > $o/${tag}_ncu_records4k.log 2>&1
python tools/ncu_summary.py $o/${tag}_k5_k1_records4k.ncu-rep $o/${tag}_ncu_k5_k1_records4k_summary.json >> $o/${tag}_ncu_summary.log 2>&1
timeout 600 python tools/bench_incremental.py --streams 4096 --pieces 32 > $o/${tag}_bench_incremental_4096x32.json 2> $o/${tag}_bench_incremental.err
PZ_TRACE=1 timeout 600 python bench.py --config huge --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --others none 2>&1 > /dev/null | grep "pz-k4" | tail -8 > $o/${tag}_huge_timeline.txt
cat $o/${tag}_ncu_summary.log; cut -c1-600 $o/${tag}_bench_incremental_4096x32.json; cat $o/${tag}_huge_timeline.txt | cut -c1-300
