"""The N>1 host logic on CPU: two gloo ranks shard a batch, each produces the verdicts of its own
range (with the oracle -- there is no GPU here), and the optional gather returns the full batch."""
import os
import socket
import sys
import zlib

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pure_zlib_b200 import shard  # noqa: E402


def _corpus():
    import streams
    rng = np.random.default_rng(8)
    out = []
    for i in range(37):
        n = int(rng.integers(0, 20_000))
        data = streams.small_text(n, i) if i % 3 else rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        z = zlib.compress(data, [1, 6, 9][i % 3])
        if i % 11 == 5:
            z = z[:-2]          # a truncated stream keeps its own verdict
        out.append(z)
    return out


def _verdicts(cases):
    from oracle import oracle
    rec = np.zeros(len(cases), dtype=shard.RESULT_DTYPE)
    for i, z in enumerate(cases):
        o = oracle.decompress(z)
        rec[i] = (o.status, o.detail, o.out_len, o.adler_computed, o.adler_stored, 0, o.payload[0], o.payload[1])
    return rec


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = _corpus()
    ranges = shard.shard_ranges([len(z) for z in cases], world)
    a, b = ranges[rank]
    local = _verdicts(cases[a:b])
    full = shard.gather_verdicts(local, ranges, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ranges, full.tobytes()))


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 100, 4096):
            lens = rng.integers(1, 100_000, n)
            r = shard.shard_ranges(lens, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            if n >= 64 * world:
                sums = [int(lens[a:b].sum()) for a, b in r]
                assert max(sums) - min(sums) <= 2 * int(lens.max())
    # homogeneous batches split evenly
    assert shard.shard_ranges([81933] * 4096, 8) == [(512 * i, 512 * (i + 1)) for i in range(8)]


def test_two_rank_gloo_gather_equals_single_run():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _verdicts(_corpus())
    for rank, ranges, raw in got:
        full = np.frombuffer(raw, dtype=shard.RESULT_DTYPE)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(want)
        assert (full == want).all(), rank
