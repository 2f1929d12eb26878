#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the batch inflate path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--others auto|none|all|a,b,c]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference algorithm on the host cores

A "step" is one pass of the hot path over one batch.  The headline line is BASELINE configs[1]
(`text256k`: 4096 x 256 KiB text, level 6); at N = 1 the same line carries `other_configs`: the other
BASELINE workloads (levels 1 and 9 of configs[1], configs[2] `records4k`, configs[3] `huge`, configs[4]
`stored16m`), each with value / ms_per_step / roofline / e2e and the same all-streams verdict + Adler-32
assertion and byte-exact spot checks as the headline.

Under torchrun every rank inflates its own copy of the configured batch (weak scaling, no collective on
the data path); with --shard the ranks split ONE batch by contiguous ranges balanced by compressed bytes
(pure_zlib_b200/shard.py; strong scaling) and --gather adds the optional NCCL gather of output slabs
and verdicts after the timed region.  `value` times the kernels with the compressed batch resident in
HBM; `e2e` goes through pz_inflate_batch_contig with pinned HOST buffers (H2D + kernels + D2H inside
the timed region); `e2e_shim` through pz_decompress_batch with pageable pointer arrays
(exactly what the Haskell shim binds).

The oracle (oracle/) is executed here only by the cpu_baseline leg and by --impl reference: it is the
CPU restatement of the reference's algorithm (the Haskell reference needs GHC < 9.2; the harness probes
for `ghc` at run time and says what it found -- DESIGN.md).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
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
    "text256k_l1": dict(workload="4096 x 256 KiB synthetic text, system zlib level 1 (BASELINE configs[1], north_star 'levels 1/6/9')", n=4096),
    "text256k_l9": dict(workload="4096 x 256 KiB synthetic text, system zlib level 9 (BASELINE configs[1], north_star 'levels 1/6/9')", n=4096),
    "records4k": dict(workload="2^20 x 4 KiB text records, 75% Z_FIXED / 25% dynamic (BASELINE configs[2])", n=1 << 20),
    "stored16m": dict(workload="512 x 16 MiB random bytes, zlib level 6 => stored blocks (BASELINE configs[4])", n=512),
    "huge": dict(workload="ONE zlib stream of --streams MiB (default 1024) of synthetic text, level 9: 16 MiB pieces from compressors "
                          "primed with the previous 32 KiB (zdict) and joined with Z_SYNC_FLUSH, so back-references cross every piece "
                          "boundary and the history is never reset (BASELINE configs[3]); block-parallel path K4", n=1024),
}
OTHERS = ["text256k_l1", "text256k_l9", "records4k", "huge", "stored16m"]
# generation + measurement cost of a sub-config on a 16-thread host (seconds, measured): a sub-config is skipped, and says so,
# when the rest of --budget does not cover it
COST_S = {"text256k_l1": 25, "text256k_l9": 40, "records4k": 75, "huge": 50, "stored16m": 90}

# the kernel(s) `roofline.kernel_ms` times (the batch without its Adler-32 pass), and what bounds them
_ISSUE = "issue-bound integer path (hot warps issue the symbol loops, one lane per stream): frac of HBM is expected to be small (DESIGN.md 3)"
ROOFLINE_KERNEL = {
    "text256k": ("pz_inflate_kernel", _ISSUE),
    "text256k_l1": ("pz_inflate_kernel", _ISSUE),
    "text256k_l9": ("pz_inflate_kernel", _ISSUE),
    "records4k": ("pz_fixed_kernel (K5: a thread per small fixed-Huffman stream, 75 % of the records) + pz_inflate_kernel (K1: the dynamic quarter)",
                  "K5 is bound by 32-byte DRAM sectors of LZ77 history (1.2 GB of 4 KiB histories in flight, ten times the L2: "
                  "traffic is ~13x the algorithmic bytes); K1 is the issue-bound hot-warp kernel (DESIGN.md 3)"),
    "stored16m": ("pz_stored_copy_kernel (+ pz_stored_probe_kernel; pz_inflate_kernel skips what K2 finished)",
                  "HBM-bound copy with the Adler-32 partial sums fused in; K3 only folds them"),
    "huge": ("K4: pz_blk_search/verify, block jobs on pz_inflate_kernel (16-bit decode), pz_blk_compact/tails/windows/resolve",
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
    ap.add_argument("--others", default="auto", help="other BASELINE workloads added to the line as other_configs: auto (all of them "
                                                     "when N = 1 and --config is the headline), none, all, or a comma-separated list")
    ap.add_argument("--budget", type=float, default=float(os.environ.get("PZ_BENCH_BUDGET_S", "420")),
                    help="seconds the other_configs may take in all (each is skipped, with a note, when it does not fit)")
    ap.add_argument("--shard", action="store_true", help="N ranks split ONE batch (strong scaling) instead of each taking a copy")
    ap.add_argument("--gather", action="store_true", help="after timing, gather output slabs and verdicts on every rank (NCCL) and check them")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-flags", type=int, default=0, help="PZ_F_* flags for the e2e call (4 = stage input, 8 = no progressive drain)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", type=int, default=64, help="streams re-checked byte-for-byte against the generator")
    return ap.parse_args()


# ---- corpus (generated once per node, shared through /dev/shm) --------------------------------
def corpus_path(name: str, n: int) -> str:
    return f"/dev/shm/pz_corpus_r2_{name}_{n}_{os.getuid()}.npz"


def ensure_corpus(name: str, n: int) -> float:
    """Generates the batch (fork pool: call this BEFORE the process initialises CUDA) and publishes it atomically under
    /dev/shm unless it is there already.  Returns the seconds it took."""
    from pure_zlib_b200 import corpus
    path = corpus_path(name, n)
    if os.path.exists(path):
        return 0.0
    t0 = time.time()
    c = getattr(corpus, name)(n)
    tmp = path + ".tmp.npz"
    np.savez(tmp, in_blob=c.in_blob, in_off=c.in_off, in_len=c.in_len, out_len=c.out_len, out_off=c.out_off,
             adler=c.adler, sha=np.frombuffer(c.sha256_in.encode(), dtype=np.uint8), name=np.frombuffer(c.name.encode(), dtype=np.uint8))
    os.replace(tmp, path)
    gen_s = time.time() - t0
    sys.stderr.write(f"[bench] generated {name} x{n} in {gen_s:.1f}s: {c.in_bytes} -> {c.out_bytes} bytes\n")
    return gen_s


def get_corpus(name: str, n: int, local_rank: int):
    """Local rank 0 generates and publishes the batch; the other ranks of the node wait for the file."""
    from pure_zlib_b200 import corpus
    path = corpus_path(name, n)
    gen_s = ensure_corpus(name, n) if local_rank == 0 else 0.0
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > 3600:
            raise SystemExit("bench.py: timed out waiting for the corpus from local rank 0")
        time.sleep(0.5)
    z = np.load(path)
    c = corpus.Corpus(bytes(z["name"]).decode(), z["in_blob"], z["in_off"], z["in_len"], z["out_len"], z["out_off"], z["adler"],
                      bytes(z["sha"]).decode())
    c.gen_seconds = gen_s
    return c


def sub_corpus(c, lo: int, hi: int):
    """Streams [lo, hi) of corpus c as a corpus of its own (offsets rebased; the blob is a view)."""
    from pure_zlib_b200 import corpus
    i0, i1 = int(c.in_off[lo]), int(c.in_off[hi])
    o0 = int(c.out_off[lo])
    s = corpus.Corpus(c.name, c.in_blob[i0:i1 + 64] if i1 + 64 <= len(c.in_blob) else np.concatenate([c.in_blob[i0:i1], np.zeros(64, np.uint8)]),
                      c.in_off[lo:hi + 1] - np.uint64(i0), c.in_len[lo:hi], c.out_len[lo:hi], c.out_off[lo:hi + 1] - np.uint64(o0),
                      c.adler[lo:hi], c.sha256_in)
    s.first = lo
    return s


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


def haskell_probe():
    """the Haskell reference with GHC -N, when a toolchain exists.  Returns a one-line statement."""
    ghc = shutil.which("ghc")
    if not ghc:
        return "not runnable in this image (no `ghc` on PATH: probed at run time with shutil.which)"
    ref = "/root/reference"
    if not os.path.isdir(ref):
        return f"`ghc` found at {ghc}, but the reference checkout ({ref}) is not on this machine"
    try:
        ver = subprocess.run([ghc, "--numeric-version"], capture_output=True, text=True, timeout=30).stdout.strip()
        out = os.path.join(ROOT, "baseline", "_ref", "hs")
        os.makedirs(out, exist_ok=True)
        r = subprocess.run([ghc, "-O2", "-threaded", "-rtsopts", f"-i{ref}/src", f"-outputdir={out}", "-o", os.path.join(out, "deflate"),
                            f"{ref}/Deflate.hs"], capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            return f"`ghc` {ver} found but the reference does not build with it (needs GHC < 9.2, bytestring < 0.11): {r.stderr.strip()[-200:]}"
        return f"built with ghc {ver} at {out}/deflate (timing it over the batch is the caller's next step)"
    except Exception as e:  # noqa: BLE001
        return f"`ghc` found at {ghc}; build attempt failed: {e}"


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


RES_DTYPE = np.dtype([("status", "<i4"), ("detail", "<i4"), ("out_len", "<u8"), ("adler_c", "<u4"), ("adler_s", "<u4"), ("bitpos", "<u8"),
                      ("p0", "<i8"), ("p1", "<i8")])


class Ctx:
    """What every measurement needs: the library, torch, the process group."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.a = torch, dist, a
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.distributed = self.world > 1
        self.L = None

    def init_cuda(self):
        torch, dist = self.torch, self.dist
        if self.distributed:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "NONE"  # stdout carries ONE line: no NCCL version banner (VERSION and WARN both print it)
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the host baseline)")
        torch.cuda.set_device(self.local_rank)
        from pure_zlib_b200 import _lib
        self.lib = _lib
        self.L = _lib.load()

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if not self.distributed:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if not self.distributed:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def verify_bytes(c, config, host, n_verify, first=0):
    """Byte-exact spot checks of a decoded batch (numpy view of the output blob) against the generator."""
    from pure_zlib_b200 import corpus as corpus_mod
    checked = 0
    if config == "huge":  # three 16 MiB pieces of the one stream, regenerated
        piece = 16 << 20
        n_pieces = (int(c.out_len[0]) + piece - 1) // piece
        for k in sorted({0, n_pieces // 2, n_pieces - 1}):
            want = corpus_mod.decoded_piece(k, min(piece, int(c.out_len[0]) - k * piece))
            o = int(c.out_off[0]) + k * piece
            assert host[o:o + len(want)].tobytes() == want, f"piece {k} differs"
            checked += 1
    else:
        for i in np.linspace(0, c.n - 1, min(n_verify, c.n)).astype(int):
            o = int(c.out_off[i])
            assert host[o:o + int(c.out_len[i])].tobytes() == corpus_mod.decoded_by_name(c.name, int(i) + first, int(c.out_len[i])), \
                f"stream {i + first} differs"
            checked += 1
    return checked


def measure(ctx: Ctx, config: str, c, steps: int, warmup: int, headline: bool, scaling: str):
    """One workload on this rank's GPU: resident-batch value, dominant-kernel roofline, end-to-end legs, checks."""
    torch, L, _lib, a = ctx.torch, ctx.L, ctx.lib, ctx.a
    p64 = C.POINTER(C.c_uint64)
    d_in = torch.from_numpy(c.in_blob).cuda()
    d_out = torch.zeros(int(c.out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    in_off = c.in_off
    batch = L.pz_batch_create(in_off.ctypes.data_as(p64), c.out_off.ctypes.data_as(p64), c.n, 0)
    batch_k1 = L.pz_batch_create(in_off.ctypes.data_as(p64), c.out_off.ctypes.data_as(p64), c.n, _lib.PZ_F_NO_ADLER)
    if not batch or not batch_k1:
        raise SystemExit("pz_batch_create failed: " + L.pz_last_error().decode())
    st = torch.cuda.current_stream().cuda_stream
    res = (_lib.PzResult * c.n)()

    def step():
        _lib.check(L.pz_batch_run(batch, d_in.data_ptr(), d_out.data_ptr(), st), "pz_batch_run")

    warmup = max(warmup, 3)
    for _ in range(warmup):
        step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank) if headline else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    # the dominant kernel alone (the batch without its Adler-32 pass), same stream, same inputs
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(steps):
        _lib.check(L.pz_batch_run(batch_k1, d_in.data_ptr(), d_out.data_ptr(), st), "pz_batch_run")
    k1.record()
    ctx.barrier()
    ms_k1 = k0.elapsed_time(k1) / steps
    # the decoder warps alone (sizing pass: same bit-stream work, no tokens, no writer warps)
    # (not for the one huge stream: its sizing pass is the serial chain K4 exists to avoid)
    ms_dec = None
    if config != "huge":
        batch_dec = L.pz_batch_create(in_off.ctypes.data_as(p64), None, c.n, _lib.PZ_F_COUNT_ONLY)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.check(L.pz_batch_run(batch_dec, d_in.data_ptr(), None, st), "pz_batch_run")
        dsteps = min(steps, 10)
        d0.record()
        for _ in range(dsteps):
            _lib.check(L.pz_batch_run(batch_dec, d_in.data_ptr(), None, st), "pz_batch_run")
        d1.record()
        ctx.barrier()
        ms_dec = d0.elapsed_time(d1) / dsteps
        L.pz_batch_destroy(batch_dec)
    clocks = sampler.stop() if sampler else None
    ms = ctx.max_over_ranks(ms)

    # verdicts of one more step: every stream must be OK with the generator's Adler-32
    step()
    _lib.check(L.pz_batch_results(batch, res, st), "pz_batch_results")
    st_arr = np.frombuffer(res, dtype=RES_DTYPE)
    checks = {}
    if not os.environ.get("PZ_BENCH_NOCHECK"):  # (only for timing experiments with kernels that skip work on purpose)
        assert (st_arr["status"] == 0).all(), f"{config}: {int((st_arr['status'] != 0).sum())} streams failed"
        assert (st_arr["out_len"] == c.out_len).all() and (st_arr["adler_c"] == c.adler).all(), f"{config}: length / Adler-32 mismatch"
        checks["all_streams_ok_and_adler32"] = int(c.n)
    host = None
    if a.verify:
        host = d_out.cpu().numpy()
        checks["streams_byte_exact_vs_generator"] = verify_bytes(c, config, host, a.verify, getattr(c, "first", 0))
    if config == "huge":
        checks["k4_done"], checks["k4_declined"] = int(L.pz_get_counter(1)), int(L.pz_get_counter(2))
        assert checks["k4_declined"] == 0, "the block-parallel path declined the stream"

    out_bytes_total = ctx.sum_over_ranks(float(c.out_bytes)) if scaling == "strong" else c.out_bytes * ctx.world
    value = out_bytes_total * steps / (ms * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    alg_bytes = c.in_bytes + c.out_bytes
    achieved = alg_bytes / (ms_k1 * 1e-3) / 1e9
    traffic_key = config if config in ("stored16m", "huge", "records4k") else "text256k" if config == "text256k" else config
    line = {"value": value, "unit": "GB/s", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "roofline": {"bound": "hbm", "kernel": ROOFLINE_KERNEL[config][0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": load_traffic(traffic_key), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_k1, "decoder_only_ms": ms_dec,
                         "note": ROOFLINE_KERNEL[config][1]},
            "gpu_launches": steps * L.pz_batch_launches(batch), "checks": checks,
            "corpus_sha256": c.sha256_in, "compressed_bytes": c.in_bytes, "decoded_bytes": c.out_bytes}
    if clocks is not None:
        line["clocks"] = clocks

    # ---- end to end through the C ABI with pinned host buffers -------------------------------
    if not a.no_e2e:
        hin = L.pz_pinned_alloc(c.in_blob.nbytes)
        hout = L.pz_pinned_alloc(int(c.out_off[-1]) + 64)
        if not hin or not hout:
            raise SystemExit("pz_pinned_alloc failed: " + L.pz_last_error().decode())
        C.memmove(hin, c.in_blob.ctypes.data, c.in_blob.nbytes)
        e2e_steps = max(3, min(steps, 10)) if headline else 3

        def e2e_step():
            _lib.check(L.pz_inflate_batch_contig(hin, in_off.ctypes.data_as(p64), hout, c.out_off.ctypes.data_as(p64), c.n, res,
                                                 None, a.e2e_flags), "pz_inflate_batch_contig")
        for _ in range(2 if headline else 1):
            e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        sec = ctx.max_over_ranks(time.perf_counter() - t0)
        hview = np.ctypeslib.as_array((C.c_uint8 * int(c.out_off[-1])).from_address(hout))
        st2 = np.frombuffer(res, dtype=RES_DTYPE)
        assert (st2["status"] == 0).all() and (st2["adler_c"] == c.adler).all(), f"{config}: e2e verdicts differ"
        verify_bytes(c, config, hview, min(a.verify, 8) or 1, getattr(c, "first", 0))
        ceiling = None
        try:  # what N GPUs of one box can copy home at all (tools/pcie_ceiling.py, measured on this pool)
            ceiling = json.load(open(os.path.join(ROOT, "profiles", "pcie_ceiling.json")))["d2h_GBps_by_gpus"].get(str(ctx.world))
        except Exception:
            pass
        line["e2e"] = {"value": out_bytes_total * e2e_steps / sec / 1e9, "unit": "GB/s", "ceiling": ceiling,
                       "ceiling_note": "aggregate pinned D2H GB/s of N GPUs on one box with the H2D copies running (profiles/pcie_ceiling.json, measured "
                                       "with tools/pcie_ceiling.py on ONE box of the pool; boxes differ by 10-20 %, so e2e can pass it): "
                                       "the host side of the box bounds e2e, not the kernels",
                       "h2d_bytes_per_step": int(c.in_off[-1]) + 3 * 8 * (c.n + 1),
                       "d2h_bytes_per_step": int(c.out_off[-1]) + 48 * c.n, "steps": e2e_steps,
                       "api": "pz_inflate_batch_contig(host pinned in/out): H2D, kernels and D2H inside the timed region"}
        del hview
        L.pz_pinned_free(hin)
        L.pz_pinned_free(hout)
        # ---- what the Haskell shim binds: pz_decompress_batch over pageable pointer arrays ----------------------
        if headline and ctx.world == 1:
            n = c.n
            pageable_in = c.in_blob.copy()   # ordinary (pageable) host memory, like the payload of a strict ByteString
            base_in = pageable_in.ctypes.data
            ptrs = (C.c_void_p * n)(*[base_in + int(c.in_off[i]) for i in range(n)])
            lens = (C.c_size_t * n)(*[int(x) for x in c.in_len])
            optrs = (C.c_void_p * n)()
            handle = C.c_void_p()

            def shim_step(keep=False):
                _lib.check(L.pz_decompress_batch(ptrs, lens, n, res, optrs, C.byref(handle), 0), "pz_decompress_batch")
                if not keep:
                    L.pz_outputs_free(handle)
            shim_step()
            t0 = time.perf_counter()
            for _ in range(3):
                shim_step()
            sec = time.perf_counter() - t0
            shim_step(keep=True)
            st3 = np.frombuffer(res, dtype=RES_DTYPE)
            assert (st3["status"] == 0).all() and (st3["adler_c"] == c.adler).all() and (st3["out_len"] == c.out_len).all()
            from pure_zlib_b200 import corpus as corpus_mod
            for i in np.linspace(0, n - 1, 8).astype(int):
                assert C.string_at(optrs[i], int(c.out_len[i])) == corpus_mod.decoded_by_name(c.name, int(i), int(c.out_len[i])), f"shim stream {i}"
            L.pz_outputs_free(handle)
            line["e2e_shim"] = {"value": c.out_bytes * 3 / sec / 1e9, "unit": "GB/s", "steps": 3,
                                "api": "pz_decompress_batch over pageable pointer arrays: the one call haskell/Codec/Compression/Zlib.hs:"
                                       "decompressBatch makes (host packing, sizing launch, decode launch, results in a library-owned pinned block)"}
            del pageable_in
    L.pz_batch_destroy(batch)
    L.pz_batch_destroy(batch_k1)
    del d_in, d_out, host
    torch.cuda.empty_cache()
    return line, st_arr.copy()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[a.config]
    n = a.streams or cfg["n"]
    scaling = "strong" if (a.shard and world > 1) else "weak"
    config = {"workload": cfg["workload"], "config": a.config, "streams_per_gpu": n,
              "l2": "per-step working set (compressed + decoded batch) exceeds the 126 MB L2; no flush needed",
              "parallelism": (f"{world} rank(s) split one batch by contiguous ranges balanced by compressed bytes (strong scaling), "
                              if scaling == "strong" else f"{world} independent shard(s), ") + "no collective on the data path"}
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
        sample = int(min(c.n, max(threads, (threads * 35e6 * 6.0) // per_stream)))  # ~6 s per step at ~35 MB/s/thread
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
                          "gpu_launches": 0, "reference_haskell": haskell_probe()}))
        return

    # ---- our arm --------------------------------------------------------------------------
    ctx = Ctx(a)
    # corpora first (fork pools) -- CUDA / NCCL are not initialised yet
    c_full = get_corpus(a.config, n, local_rank)
    want = a.others
    if want == "auto":
        want = "all" if (world == 1 and a.config == "text256k" and not a.streams) else "none"
    names = [] if (want == "none" or world > 1) else OTHERS if want == "all" else [x for x in want.split(",") if x]
    t_begin = time.time()
    others = {}
    gen_seconds = {}
    for name in names:
        left = a.budget - (time.time() - t_begin)
        need = COST_S.get(name, 60) * (0.4 if os.path.exists(corpus_path(name, CONFIGS[name]["n"])) else 1.0)
        if left < need:
            others[name] = {"skipped": f"{left:.0f} s of --budget {a.budget:.0f} s left, about {need:.0f} s needed; run "
                                       f"`python bench.py --config {name}` for this workload alone"}
            continue
        gen_seconds[name] = ensure_corpus(name, CONFIGS[name]["n"])
    ctx.init_cuda()
    c = c_full
    ranges = None
    if scaling == "strong":
        from pure_zlib_b200 import shard
        ranges = shard.shard_ranges(c_full.in_len, world)
        lo, hi = ranges[rank]
        c = sub_corpus(c_full, lo, hi)
        config["streams_per_gpu"] = f"{c_full.n} in all, contiguous ranges of about {c_full.n // world}"
    line, verdicts = measure(ctx, a.config, c, a.steps, a.warmup, True, scaling)
    head = {"metric": METRIC, "value": line.pop("value"), "unit": line.pop("unit"), "n_gpus": world, "steps": line.pop("steps"),
            "warmup": line.pop("warmup"), "ms_per_step": line.pop("ms_per_step"), "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config}
    head.update(line)
    line = head

    # ---- optional gather of the results (NCCL, outside the timed region) ------------------------
    if a.gather and ctx.distributed:
        from pure_zlib_b200 import shard
        torch = ctx.torch
        if ranges is None:  # weak scaling: every rank holds the same batch; gather the verdict records only
            ranges = [(0, c.n)] * world
        t0 = time.perf_counter()
        rec = np.zeros(len(verdicts), dtype=shard.RESULT_DTYPE)
        rec.view(np.uint8)[:] = verdicts.view(np.uint8)
        if scaling == "strong":
            full = shard.gather_verdicts(rec, ranges, rank, world, device=torch.device("cuda", ctx.local_rank))
            assert (full["status"] == 0).all() and (full["adler_computed"] == c_full.adler).all() and (full["out_len"] == c_full.out_len).all()
            # output slabs: decode once more into a device buffer and all-gather the slabs (padded to the longest)
            p64 = C.POINTER(C.c_uint64)
            d_in = torch.from_numpy(c.in_blob).cuda()
            d_out = torch.zeros(int(c.out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
            res = (ctx.lib.PzResult * c.n)()
            ctx.lib.check(ctx.L.pz_inflate_batch_contig(d_in.data_ptr(), c.in_off.ctypes.data_as(p64), d_out.data_ptr(), c.out_off.ctypes.data_as(p64),
                                                        c.n, res, None, 0), "pz_inflate_batch_contig")
            slabs = shard.gather_outputs(d_out[: int(c.out_off[-1])], [int(c_full.out_off[b] - c_full.out_off[a_]) for a_, b in ranges], rank, world)
            got = slabs.cpu().numpy()
            checked = verify_bytes(c_full, a.config, got, 16)
            line["gather"] = {"verdicts": int(len(full)), "output_bytes": int(got.nbytes), "streams_byte_exact_vs_generator": checked,
                              "seconds": time.perf_counter() - t0, "collective": "torch.distributed all_gather over NCCL, outside the timed region"}
        else:
            full = shard.gather_verdicts(rec, [(r * c.n, (r + 1) * c.n) for r in range(world)], rank, world, device=torch.device("cuda", ctx.local_rank))
            assert (full["status"] == 0).all() and len(full) == c.n * world
            line["gather"] = {"verdicts": int(len(full)), "seconds": time.perf_counter() - t0,
                              "collective": "torch.distributed all_gather over NCCL, outside the timed region"}

    # ---- the reference algorithm on this box's host cores (rank 0, N=1 only) ------------------
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        per_stream = float(c.out_len.mean())
        sample = int(min(c.n, max(threads, (threads * 35e6 * 12.0) // per_stream)))
        b, s = cpu_reference(c, np.arange(sample), threads)
        line["cpu_baseline"] = {"value": b / s / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                "sample": f"first {sample} of {c.n} streams, oracle/pz_oracle.c (restatement of pure-zlib; "
                                          "the Haskell reference needs GHC, absent here)", "seconds": s,
                                "reference_haskell": haskell_probe()}
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

    # ---- the other BASELINE workloads (N = 1 only; their corpora were generated before CUDA came up) -------
    if names:
        del c, c_full
        for name in names:
            if name in others:
                continue
            t0 = time.time()
            try:
                oc = get_corpus(name, CONFIGS[name]["n"], 0)
                sub, _ = measure(ctx, name, oc, max(3, min(a.steps, 5)), 3, False, "weak")
                sub["config"] = {"workload": CONFIGS[name]["workload"], "config": name, "streams_per_gpu": 1 if name == "huge" else oc.n}
                sub["seconds"] = {"corpus": gen_seconds.get(name, 0.0), "measure": time.time() - t0}
                others[name] = sub
                del oc
            except AssertionError as e:  # a failed check is reported, never hidden: the headline stays valid
                others[name] = {"failed": str(e) or "assertion"}
        line["other_configs"] = {k: others[k] for k in names}
    if rank == 0:
        print(json.dumps(line))
    if ctx.distributed:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
