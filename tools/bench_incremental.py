#!/usr/bin/env python
"""Throughput of the incremental API (pz_stream_feed / pz_stream_pump / pz_stream_next): N concurrent
`decompressIncremental` consumers, each fed its stream in K pieces, one pump (one kernel launch) per round.
Prints one JSON line: decompressed GB/s end to end (host chunks in, host chunks out), per-round times, the device
and pinned host memory the contexts hold at their peak (they are O(1) in the stream: live input, the last 32 KiB of
output plus room for one launch, a checkpoint), and the same run with one pz_stream_feed per stream instead of one
pz_stream_feed_many per round.

  python tools/bench_incremental.py --streams 1024 --pieces 8
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pure_zlib_b200 import _lib, corpus  # noqa: E402


def run(L, c, pieces: int, many: bool):
    n = c.n
    streams = [L.pz_stream_new() for _ in range(n)]
    assert all(streams)
    arr = (C.c_void_p * n)(*streams)
    blobs = [bytes(c.in_blob[int(c.in_off[i]): int(c.in_off[i]) + int(c.in_len[i])]) for i in range(n)]
    steps = [(len(z) + pieces - 1) // pieces for z in blobs]
    ptr, ln, res = C.c_void_p(), C.c_size_t(), _lib.PzResult()
    out_bytes = 0
    rounds = []
    t0 = time.perf_counter()
    for r in range(pieces):
        t1 = time.perf_counter()
        if many:  # one call, one copy across the bus for the whole round
            parts = [blobs[i][r * steps[i]:(r + 1) * steps[i]] for i in range(n)]
            _lib.check(L.pz_stream_feed_many(arr, (C.c_char_p * n)(*parts), (C.c_size_t * n)(*[len(p) for p in parts]), n), "feed_many")
        else:
            for i in range(n):
                piece = blobs[i][r * steps[i]:(r + 1) * steps[i]]
                _lib.check(L.pz_stream_feed(streams[i], piece, len(piece)), "feed")
        t2 = time.perf_counter()
        _lib.check(L.pz_stream_pump(arr, n), "pump")
        t3 = time.perf_counter()
        done = 0
        for i in range(n):
            while True:
                ev = L.pz_stream_next(streams[i], C.byref(ptr), C.byref(ln), C.byref(res))
                if ev != _lib.PZ_S_CHUNK:
                    break
                out_bytes += ln.value
            done += ev == _lib.PZ_S_DONE
        t4 = time.perf_counter()
        rounds.append({"feed_ms": (t2 - t1) * 1e3, "pump_ms": (t3 - t2) * 1e3, "drain_ms": (t4 - t3) * 1e3})
    total = time.perf_counter() - t0
    assert done == n, (done, n)
    assert out_bytes == int(c.out_len.sum()), (out_bytes, int(c.out_len.sum()))
    resumed = sum(L.pz_stream_counter(s, _lib.PZ_SC_RESUMED) for s in streams)
    mem = {"device_peak_bytes_per_context": max(L.pz_stream_counter(s, _lib.PZ_SC_DEVICE_PEAK) for s in streams),
           "device_peak_bytes_all_contexts": sum(L.pz_stream_counter(s, _lib.PZ_SC_DEVICE_PEAK) for s in streams),
           "host_bytes_all_contexts": sum(L.pz_stream_counter(s, _lib.PZ_SC_HOST_BYTES) for s in streams)}
    for s in streams:
        L.pz_stream_free(s)
    return total, rounds, out_bytes, resumed, mem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--pieces", type=int, default=8)
    a = ap.parse_args()
    L = _lib.load()
    c = corpus.text256k(a.streams, workers=min(16, os.cpu_count() or 1))
    line = {"metric": "decompressed GB/s through the incremental API", "unit": "GB/s", "streams": a.streams, "pieces": a.pieces,
            "workload": f"{a.streams} x 256 KiB synthetic text (level 6), each stream fed in {a.pieces} pieces, one pz_stream_pump per round"}
    for label, many in (("warmup", True), ("resumed", True), ("resumed_single_feeds", False)):
        total, rounds, out_bytes, resumed, mem = run(L, c, a.pieces, many)
        if label == "warmup":
            continue
        line[label] = {"value": out_bytes / total / 1e9, "seconds": total, "pump_ms": [round(r["pump_ms"], 2) for r in rounds],
                       "feed_ms": round(sum(r["feed_ms"] for r in rounds), 1), "drain_ms": round(sum(r["drain_ms"] for r in rounds), 1),
                       "pump_ms_total": round(sum(r["pump_ms"] for r in rounds), 1), "launches_from_checkpoint": int(resumed), "memory": mem}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
