#!/bin/bash
# round 2, GPU call E (2 GPUs): whole suite (incl. the in-library multi-GPU test), one-process multi-GPU bench
o=gpurun_out; tag=r02e
( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $o/${tag}_pytest_gpu.log ) 2> $o/${tag}_pytest.time
timeout 900 python tools/bench_inlib_multigpu.py --config text256k --gpus 1,2 > $o/${tag}_inlib_text256k.jsonl 2> $o/${tag}_inlib.err
timeout 900 python tools/bench_inlib_multigpu.py --config records4k --gpus 1,2 > $o/${tag}_inlib_records4k.jsonl 2>> $o/${tag}_inlib.err
tail -6 $o/${tag}_pytest_gpu.log; cat $o/${tag}_pytest.time; cat $o/${tag}_inlib_text256k.jsonl $o/${tag}_inlib_records4k.jsonl | cut -c1-300; tail -n 3 $o/${tag}_inlib.err
