#!/bin/bash
# round 2, GPU call V: what K5's history reads and its stores cost (timing experiments with wrong output: -DPZ_EXP_K5_NOLOAD / _NOSTORE)
o=gpurun_out; tag=r02v
for v in base noload nostore; do
  lib=pure_zlib_b200/libpzcuda_$v.so; [ $v = base ] && lib=pure_zlib_b200/libpzcuda.so
  PZ_BENCH_NOCHECK=1 PZ_LIBPZCUDA=$PWD/$lib ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pz_fixed_kernel -s 3 -c 1 --csv --log-file $o/${tag}_k5_$v.csv python bench.py --steps 1 --warmup 3 --config records4k --others none --no-e2e --no-cpu-baseline --verify 0 > $o/${tag}_$v.log 2>&1
  grep "pz_fixed_kernel" $o/${tag}_k5_$v.csv | awk -F'","' '{print "'$v'", $(NF-2), $(NF-1), $NF}'
done
