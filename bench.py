#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the batch inflate path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config text256k|records4k|stored16m]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference algorithm on the host cores

A "step" is one pass of the hot path over one batch: every rank inflates its own copy of the
configured batch (weak scaling, no collective on the data path).  `value` times the kernels
with the compressed batch resident in HBM; `e2e` goes through pz_inflate_batch_contig with
pinned HOST buffers (H2D + kernels + D2H inside the timed region).

The oracle (oracle/) is executed here only by the cpu_baseline leg and by --impl reference:
it is the CPU restatement of the reference's algorithm (no GHC exists in this image, so the
Haskell reference itself cannot be timed -- DESIGN.md).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decompressed GB/s, 256 KiB-stream batch"
CONFIGS = {
    "text256k": dict(workload="4096 x 256 KiB synthetic text, system zlib level 6 (BASELINE configs[1])", n=4096),
    "records4k": dict(workload="2^20 x 4 KiB text records, 75% Z_FIXED / 25% dynamic (BASELINE configs[2])", n=1 << 20),
    "stored16m": dict(workload="512 x 16 MiB random bytes, zlib level 6 => stored blocks (BASELINE configs[4])", n=512),
    "huge": dict(workload="ONE zlib stream of --streams MiB (default 1024) of synthetic text, level 9, compressed in 16 MiB pieces "
                          "and stitched (BASELINE configs[3]); block-parallel path K4", n=1024),
}


# the kernel(s) `roofline.kernel_ms` times (the batch without its Adler-32 pass), and what bounds them
_ISSUE = "issue-bound integer path (one warp issues the symbol loops of 28 streams): frac of HBM is expected to be small (DESIGN.md 3)"
ROOFLINE_KERNEL = {
    "text256k": ("pz_inflate_kernel", _ISSUE),
    "records4k": ("pz_inflate_kernel", _ISSUE),
    "stored16m": ("pz_stored_copy_kernel (+ pz_stored_probe_kernel; pz_inflate_kernel skips what K2 finished)",
                  "HBM-bound copy with the Adler-32 partial sums fused in; K3 only folds them"),
    "huge": ("K4: pz_blk_search/verify, block jobs on pz_inflate_kernel (sizing + 16-bit decode), pz_blk_tails/windows/resolve",
             "one stream: block-parallel decode with a 2-byte symbol buffer; frac of HBM is small"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="text256k", choices=list(CONFIGS))
    ap.add_argument("--streams", type=int, default=0, help="override the number of streams (debug only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-flags", type=int, default=0, help="PZ_F_* flags for the e2e call (4 = stage input, 8 = no progressive drain)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", type=int, default=64, help="streams re-checked byte-for-byte against the generator")
    return ap.parse_args()


# ---- corpus (generated once per node, shared through /dev/shm) --------------------------------
def get_corpus(name: str, n: int, local_rank: int):
    """Local rank 0 generates (fork pool, before any CUDA/NCCL initialisation) and publishes the
    batch atomically under /dev/shm; the other ranks of the node wait for the file."""
    from pure_zlib_b200 import corpus
    path = f"/dev/shm/pz_corpus_{name}_{n}_{os.getuid()}.npz"
    if local_rank == 0 and not os.path.exists(path):
        t0 = time.time()
        c = getattr(corpus, name)(n)
        tmp = path + ".tmp.npz"
        np.savez(tmp, in_blob=c.in_blob, in_off=c.in_off, in_len=c.in_len, out_len=c.out_len, out_off=c.out_off,
                 adler=c.adler, sha=np.frombuffer(c.sha256_in.encode(), dtype=np.uint8))
        os.replace(tmp, path)
        sys.stderr.write(f"[bench] generated {name} x{n} in {time.time() - t0:.1f}s: {c.in_bytes} -> {c.out_bytes} bytes\n")
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > 3600:
            raise SystemExit("bench.py: timed out waiting for the corpus from local rank 0")
        time.sleep(0.5)
    z = np.load(path)
    return corpus.Corpus(name, z["in_blob"], z["in_off"], z["in_len"], z["out_len"], z["out_off"], z["adler"],
                         bytes(z["sha"]).decode())


# ---- the reference algorithm on the host cores -------------------------------------------------
def cpu_reference(c, idxs, threads):
    """Decodes streams `idxs` of corpus `c` with the oracle on `threads` host threads (ctypes
    releases the GIL).  Returns (decompressed bytes, seconds)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle
    L = oracle.lib()
    base = c.in_blob.ctypes.data
    one = (C.c_size_t * 1)

    def work(chunk):
        cap = int(c.out_len[chunk].max()) if len(chunk) else 0
        out = C.create_string_buffer(max(cap, 1))
        res = oracle.PzoResult()
        n_ev = C.c_size_t(0)
        pub = C.c_uint64(0)
        total = 0
        for i in chunk:
            ln = one(int(c.in_len[i]))
            L.pzo_decompress(C.c_void_p(base + int(c.in_off[i])), ln, 1, out, cap, C.byref(res), None, 0, C.byref(n_ev),
                             C.byref(pub))
            assert res.status == 0 and res.adler_computed == int(c.adler[i])
            total += int(res.out_len)
        return total

    chunks = [idxs[k::threads] for k in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        total = sum(ex.map(work, chunks))
    return total, time.perf_counter() - t0


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(config):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[config]["dram_bytes_per_launch"]
    except Exception:
        return None


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[a.config]
    n = a.streams or cfg["n"]
    config = {"workload": cfg["workload"], "config": a.config, "streams_per_gpu": n,
              "l2": "per-step working set (compressed + decoded batch) exceeds the 126 MB L2; no flush needed",
              "parallelism": f"{world} independent shard(s), no collective on the data path"}
    if a.config == "huge":
        config["streams_per_gpu"], config["stream_mib"] = 1, n
    elif a.streams:
        config["workload"] += f" [DEBUG: {n} streams]"

    if a.impl == "reference":
        if rank != 0:
            return
        c = get_corpus(a.config, n, 0)
        threads = os.cpu_count() or 1
        per_stream = float(c.out_len.mean())
        sample = int(min(n, max(threads, (threads * 35e6 * 6.0) // per_stream)))  # ~6 s per step at ~35 MB/s/thread
        idxs = np.arange(sample)
        for _ in range(min(a.warmup, 1)):
            cpu_reference(c, idxs[: max(threads, sample // 8)], threads)
        tot, sec = 0, 0.0
        for _ in range(a.steps):
            b, s = cpu_reference(c, idxs, threads)
            tot += b; sec += s
        v = tot / sec / 1e9
        kind = "port"
        note = f"oracle restatement of pure-zlib's decoder on {threads} host threads; first {sample} streams of the batch per step"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": sec / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "GB/s", "cores": threads, "kind": kind, "sample": note},
                          "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0, "reference_haskell": "not runnable in this image (no GHC)"}))
        return

    # ---- our arm --------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")

    # corpus first (fork pool) -- CUDA / NCCL are not initialised yet
    c = get_corpus(a.config, n, local_rank)
    if distributed:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    from pure_zlib_b200 import _lib
    L = _lib.load()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    d_in = torch.from_numpy(c.in_blob).cuda()
    d_out = torch.zeros(int(c.out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    p64 = C.POINTER(C.c_uint64)
    # streams end where their bytes end, not at the alignment padding
    in_off = c.in_off
    batch = L.pz_batch_create(in_off.ctypes.data_as(p64), c.out_off.ctypes.data_as(p64), c.n, 0)
    batch_k1 = L.pz_batch_create(in_off.ctypes.data_as(p64), c.out_off.ctypes.data_as(p64), c.n, _lib.PZ_F_NO_ADLER)
    if not batch or not batch_k1:
        raise SystemExit("pz_batch_create failed: " + L.pz_last_error().decode())
    st = torch.cuda.current_stream().cuda_stream
    res = (_lib.PzResult * c.n)()

    def step():
        _lib.check(L.pz_batch_run(batch, d_in.data_ptr(), d_out.data_ptr(), st), "pz_batch_run")

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # the dominant kernel alone (K1 inflate), same stream, same inputs
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(a.steps):
        _lib.check(L.pz_batch_run(batch_k1, d_in.data_ptr(), d_out.data_ptr(), st), "pz_batch_run")
    k1.record()
    barrier()
    # the decoder warps alone (sizing pass: same bit-stream work, no tokens, no writer warps)
    # (not for the one huge stream: its sizing pass is the serial chain K4 exists to avoid)
    batch_dec = None if a.config == "huge" else L.pz_batch_create(in_off.ctypes.data_as(p64), None, c.n, _lib.PZ_F_COUNT_ONLY)
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_dec = None
    if batch_dec:
        _lib.check(L.pz_batch_run(batch_dec, d_in.data_ptr(), None, st), "pz_batch_run")
        d0.record()
        for _ in range(a.steps):
            _lib.check(L.pz_batch_run(batch_dec, d_in.data_ptr(), None, st), "pz_batch_run")
        d1.record()
        barrier()
        ms_dec = d0.elapsed_time(d1) / a.steps
        L.pz_batch_destroy(batch_dec)
    clocks = sampler.stop()
    ms_k1 = k0.elapsed_time(k1) / a.steps
    if distributed:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # verdicts of the last step: every stream must be OK with the generator's Adler-32
    step()
    _lib.check(L.pz_batch_results(batch, res, st), "pz_batch_results")
    st_arr = np.frombuffer(res, dtype=np.dtype([("status", "<i4"), ("detail", "<i4"), ("out_len", "<u8"), ("adler_c", "<u4"),
                                                ("adler_s", "<u4"), ("bitpos", "<u8"), ("p0", "<i8"), ("p1", "<i8")]))
    if not os.environ.get("PZ_BENCH_NOCHECK"):  # (only for timing experiments with kernels that skip work on purpose)
        assert (st_arr["status"] == 0).all(), f"{int((st_arr['status'] != 0).sum())} streams failed"
        assert (st_arr["out_len"] == c.out_len).all() and (st_arr["adler_c"] == c.adler).all()
    if a.verify:
        from pure_zlib_b200 import corpus as corpus_mod
        host = d_out.cpu().numpy()
        if a.config == "huge":  # three 16 MiB pieces of the one stream, regenerated
            piece = 16 << 20
            n_pieces = (int(c.out_len[0]) + piece - 1) // piece
            for k in sorted({0, n_pieces // 2, n_pieces - 1}):
                want = corpus_mod.decoded_piece(k, min(piece, int(c.out_len[0]) - k * piece))
                o = int(c.out_off[0]) + k * piece
                assert host[o:o + len(want)].tobytes() == want, f"piece {k} differs"
        else:
            for i in np.linspace(0, c.n - 1, min(a.verify, c.n)).astype(int):
                o = int(c.out_off[i])
                assert host[o:o + int(c.out_len[i])].tobytes() == corpus_mod.decoded(c, int(i)), f"stream {i} differs"
        del host

    out_bytes_total = c.out_bytes * world
    value = out_bytes_total * a.steps / (ms * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    alg_bytes = c.in_bytes + c.out_bytes
    achieved = alg_bytes / (ms_k1 * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "kernel": ROOFLINE_KERNEL[a.config][0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": load_traffic(a.config), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_k1, "decoder_only_ms": ms_dec,
                         "note": ROOFLINE_KERNEL[a.config][1]},
            "clocks": clocks, "gpu_launches": a.steps * L.pz_batch_launches(batch),
            "corpus_sha256": c.sha256_in, "compressed_bytes": c.in_bytes, "decoded_bytes": c.out_bytes}

    # ---- end to end through the C ABI with pinned host buffers -------------------------------
    if not a.no_e2e:
        hin = L.pz_pinned_alloc(c.in_blob.nbytes)
        hout = L.pz_pinned_alloc(int(c.out_off[-1]) + 64)
        if not hin or not hout:
            raise SystemExit("pz_pinned_alloc failed: " + L.pz_last_error().decode())
        C.memmove(hin, c.in_blob.ctypes.data, c.in_blob.nbytes)
        e2e_steps = max(3, min(a.steps, 10))

        def e2e_step():
            _lib.check(L.pz_inflate_batch_contig(hin, in_off.ctypes.data_as(p64), hout, c.out_off.ctypes.data_as(p64), c.n, res,
                                                 None, a.e2e_flags), "pz_inflate_batch_contig")
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if distributed:
            t = torch.tensor([sec], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        hview = np.ctypeslib.as_array((C.c_uint8 * int(c.out_off[-1])).from_address(hout))
        i = c.n // 2
        o = int(c.out_off[i])
        from pure_zlib_b200 import corpus as corpus_mod
        if a.config == "huge":
            assert hview[o:o + (1 << 20)].tobytes() == corpus_mod.decoded_piece(0)[: 1 << 20]
        else:
            assert hview[o:o + int(c.out_len[i])].tobytes() == corpus_mod.decoded(c, i)
        line["e2e"] = {"value": out_bytes_total * e2e_steps / sec / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": int(c.in_off[-1]) + 3 * 8 * (c.n + 1),
                       "d2h_bytes_per_step": int(c.out_off[-1]) + 48 * c.n, "steps": e2e_steps,
                       "api": "pz_inflate_batch_contig(host pinned in/out): one launch, input copied in pieces while the kernel runs, "
                              "finished column blocks of the output copied home during the decode"}
        L.pz_pinned_free(hin)
        L.pz_pinned_free(hout)

    # ---- the reference algorithm on this box's host cores (rank 0, N=1 only) ------------------
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        per_stream = float(c.out_len.mean())
        sample = int(min(c.n, max(threads, (threads * 35e6 * 12.0) // per_stream)))
        b, s = cpu_reference(c, np.arange(sample), threads)
        line["cpu_baseline"] = {"value": b / s / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                "sample": f"first {sample} of {c.n} streams, oracle/pz_oracle.c (restatement of pure-zlib; "
                                          "the Haskell reference needs GHC, absent here)", "seconds": s}
        # second stand-in: the host's system zlib on the same sample and threads -- NOT the reference
        # (whose README, lines 6-8, puts pure-zlib "roughly 100x" behind it); zlib.decompress releases the GIL
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        blobs = [bytes(c.in_blob[int(c.in_off[i]): int(c.in_off[i]) + int(c.in_len[i])]) for i in range(sample)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            zb = sum(ex.map(lambda part: sum(len(zlib.decompress(z)) for z in part), [blobs[k::threads] for k in range(threads)]))
        line["cpu_baseline"]["host_zlib"] = {"value": zb / (time.perf_counter() - t0) / 1e9, "unit": "GB/s", "cores": threads,
                                             "note": "system zlib inflate on the same sample; context only, not the reference"}
    L.pz_batch_destroy(batch)
    L.pz_batch_destroy(batch_k1)
    if rank == 0:
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
