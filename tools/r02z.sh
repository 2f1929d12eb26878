#!/bin/bash
# round 2, GPU call Z: the big-stream fuzz with the block-parallel path K4 taking every stream of >= 16 KiB compressed
o=gpurun_out; tag=r02z
timeout 1500 python tools/fuzz_gpu.py --big --huge-bytes 16384 --seeds 24 --per-seed 300 --first-seed 700 --incremental 0 > $o/${tag}_fuzz_k4.json 2> $o/${tag}_fuzz.err; echo "rc=$?"
cut -c1-1200 $o/${tag}_fuzz_k4.json; tail -3 $o/${tag}_fuzz.err
