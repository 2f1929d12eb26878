#!/bin/bash
# round 2, GPU call Q: phase clocks of K1's service groups on small streams
o=gpurun_out; tag=r02q
export PZ_LIBPZCUDA=$PWD/pure_zlib_b200/libpzcuda_phases.so
PZ_NO_K6=1 timeout 600 python tools/phase_probe.py records4k 1048576 > $o/${tag}_phases_records4k.log 2>&1
PZ_NO_K6=1 PZ_NO_K5=1 timeout 600 python tools/phase_probe.py records4k 262144 > $o/${tag}_phases_records4k_nok5.log 2>&1
timeout 600 python tools/phase_probe.py text256k 4096 > $o/${tag}_phases_text256k.log 2>&1
tail -n 20 $o/${tag}_phases_*.log
