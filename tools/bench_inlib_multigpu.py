#!/usr/bin/env python
"""ONE process, N devices: pz_inflate_batch_contig on pinned host blobs with
pz_config.n_devices = N -- the library shards the batch over the devices (a worker thread, pinned staging and a
CUDA stream set per device).  Prints one JSON line per N: end-to-end decompressed GB/s (H2D + kernels + D2H inside
the timed region) of the SAME batch (strong scaling).  Each N runs in a child process (the library initialises once).

  python tools/bench_inlib_multigpu.py --config records4k --gpus 1,2,4,8
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(config, n_dev, steps):
    import numpy as np
    import bench
    c = bench.get_corpus(config, bench.CONFIGS[config]["n"], 0)
    from pure_zlib_b200 import _lib
    L = _lib.load()
    cfg = _lib.PzConfig()
    cfg.device, cfg.n_devices = -1, n_dev
    for k in range(n_dev):
        cfg.devices[k] = k
    _lib.check(L.pz_init(C.byref(cfg)), "pz_init")
    assert L.pz_device_count() == n_dev
    p64 = C.POINTER(C.c_uint64)
    hin = L.pz_pinned_alloc(c.in_blob.nbytes)
    hout = L.pz_pinned_alloc(int(c.out_off[-1]) + 64)
    C.memmove(hin, c.in_blob.ctypes.data, c.in_blob.nbytes)
    res = (_lib.PzResult * c.n)()

    def step():
        _lib.check(L.pz_inflate_batch_contig(hin, c.in_off.ctypes.data_as(p64), hout, c.out_off.ctypes.data_as(p64), c.n, res, None, 0), "contig")
    step(); step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    sec = time.perf_counter() - t0
    st = np.frombuffer(res, dtype=bench.RES_DTYPE)
    assert (st["status"] == 0).all() and (st["adler_c"] == c.adler).all() and (st["out_len"] == c.out_len).all()
    hview = np.ctypeslib.as_array((C.c_uint8 * int(c.out_off[-1])).from_address(hout))
    bench.verify_bytes(c, config, hview, 16)
    print(json.dumps({"config": config, "devices_in_one_process": n_dev, "e2e_GBps": c.out_bytes * steps / sec / 1e9, "ms_per_step": sec / steps * 1e3,
                      "steps": steps, "streams": int(c.n), "scaling": "strong (one batch, contiguous ranges balanced by compressed bytes)",
                      "api": "pz_inflate_batch_contig(host pinned blobs), pz_config.n_devices = N"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="records4k")
    ap.add_argument("--gpus", default="1,2")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--child", type=int, default=0)
    a = ap.parse_args()
    if a.child:
        return child(a.config, a.child, a.steps)
    import bench
    bench.ensure_corpus(a.config, bench.CONFIGS[a.config]["n"])
    for n in [int(x) for x in a.gpus.split(",")]:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", a.config, "--child", str(n), "--steps", str(a.steps)],
                           capture_output=True, text=True, timeout=1800)
        sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"devices_in_one_process": n, "failed": r.stderr[-400:]}) + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
