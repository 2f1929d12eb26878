#!/bin/bash
# round 2, GPU call X: deep differential fuzz against the oracle (150 000 cases, three framings, two entry points)
o=gpurun_out; tag=r02x
timeout 1500 python tools/fuzz_gpu.py --seeds 60 --per-seed 2500 --first-seed 5000 --incremental 10000 > $o/${tag}_fuzz.json 2> $o/${tag}_fuzz.err; echo "rc=$?"
cut -c1-1500 $o/${tag}_fuzz.json; tail -3 $o/${tag}_fuzz.err
