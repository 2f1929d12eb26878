#!/bin/bash
# round 2, GPU call C: the whole suite after gzip / O(1) contexts / multi-device, memcheck over the new paths
o=gpurun_out; tag=r02c
( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $o/${tag}_pytest_gpu.log ) 2> $o/${tag}_pytest.time
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "gzip_and_raw or crc32 or incremental_event or multichunk or stream_pump_many or decompress_batch_single" 2>&1 | tail -12 > $o/${tag}_memcheck.log
timeout 300 python tools/bench_incremental.py --streams 1024 --pieces 8 > $o/${tag}_bench_incremental_1024x8.json 2> $o/${tag}_bench_incremental.err
cat $o/${tag}_pytest_gpu.log | tail -15; cat $o/${tag}_pytest.time; tail -5 $o/${tag}_memcheck.log; cut -c1-600 $o/${tag}_bench_incremental_1024x8.json; tail -3 $o/${tag}_bench_incremental.err
