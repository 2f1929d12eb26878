#!/usr/bin/env python
"""What the host side of ONE box can move: N GPUs each copying config 2's bytes per step --
0.317 GB host->device and 1.074 GB device->host, plain cudaMemcpyAsync on pinned buffers, both directions at once on
two CUDA streams per GPU -- with all N running concurrently.  The aggregate device->host rate is the ceiling of
`e2e` (decompressed GB/s through host buffers) at N GPUs: e2e = decoded bytes / max(copy time, kernel time).

  python tools/pcie_ceiling.py --gpus 1,2,4,8          (one process, one thread per GPU)
"""
import argparse
import json
import threading
import time

import torch

H2D_BYTES = 316_721_304
D2H_BYTES = 1_073_938_432


def worker(dev, steps, barrier, out):
    torch.cuda.set_device(dev)
    hin = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    hout = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    din = torch.empty(H2D_BYTES, dtype=torch.uint8, device=f"cuda:{dev}")
    dout = torch.empty(D2H_BYTES, dtype=torch.uint8, device=f"cuda:{dev}")
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def step(both=True, d2h=True):
        if both or not d2h:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if both or d2h:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    for mode in ("both", "d2h", "h2d"):
        step(); torch.cuda.synchronize(dev)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            step(mode == "both", mode == "d2h")
        torch.cuda.synchronize(dev)
        out[(dev, mode)] = time.perf_counter() - t0
        barrier.wait()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    avail = torch.cuda.device_count()
    for n in [int(x) for x in a.gpus.split(",")]:
        if n > avail:
            print(json.dumps({"gpus": n, "skipped": f"only {avail} visible"}))
            continue
        out = {}
        barrier = threading.Barrier(n)
        th = [threading.Thread(target=worker, args=(d, a.steps, barrier, out)) for d in range(n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        line = {"gpus": n, "steps": a.steps, "h2d_bytes_per_step": H2D_BYTES, "d2h_bytes_per_step": D2H_BYTES}
        for mode, nbytes in (("both", D2H_BYTES), ("d2h", D2H_BYTES), ("h2d", H2D_BYTES)):
            sec = max(out[(d, mode)] for d in range(n))
            key = {"both": "d2h_GBps_with_h2d_running (= e2e ceiling, decompressed GB/s)", "d2h": "d2h_GBps_alone", "h2d": "h2d_GBps_alone"}[mode]
            line[key] = n * nbytes * a.steps / sec / 1e9
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
