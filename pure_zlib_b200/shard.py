"""Multi-GPU sharding of a batch: streams are independent, so a batch is cut into
contiguous ranges, one per rank, balanced by compressed bytes.  There is no collective on the data
path; `gather_verdicts` and `gather_outputs` are the optional gather of the 48-byte verdict records and of
the decoded output slabs afterwards (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

RESULT_DTYPE = np.dtype([("status", "<i4"), ("detail", "<i4"), ("out_len", "<u8"), ("adler_computed", "<u4"),
                         ("adler_stored", "<u4"), ("err_bitpos", "<u8"), ("payload0", "<i8"), ("payload1", "<i8")])


def shard_ranges(in_len, world: int):
    """[(first, last)) per rank: contiguous, covering 0..n, sizes balanced by sum(in_len)."""
    in_len = np.asarray(in_len, dtype=np.uint64)
    n = len(in_len)
    if world <= 0:
        raise ValueError("world must be positive")
    cum = np.concatenate([[0], np.cumsum(in_len, dtype=np.float64)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_verdicts(local: np.ndarray, ranges, rank: int, world: int, device=None):
    """All ranks contribute the verdict records of their range; every rank returns the full array
    (n records).  Uses the default process group (torch.distributed must be initialised)."""
    import torch
    import torch.distributed as dist
    assert local.dtype == RESULT_DTYPE and len(local) == ranges[rank][1] - ranges[rank][0]
    longest = max(b - a for a, b in ranges)
    buf = np.zeros(longest * RESULT_DTYPE.itemsize, dtype=np.uint8)
    buf[: local.nbytes] = local.view(np.uint8)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.zeros(ranges[-1][1], dtype=RESULT_DTYPE)
    for r, (a, b) in enumerate(ranges):
        out[a:b] = parts[r].cpu().numpy()[: (b - a) * RESULT_DTYPE.itemsize].view(RESULT_DTYPE)
    return out


def gather_outputs(local_slab, slab_bytes, rank: int, world: int):
    """All ranks contribute their decoded output slab (a 1-D uint8 torch tensor, on the GPU under NCCL or on the
    host under gloo; rank r's slab has slab_bytes[r] bytes); every rank returns the concatenation in rank order,
    i.e. the output blob of the whole batch.  Slabs are padded to the longest for the collective (all_gather
    needs equal shapes) and trimmed afterwards.  Uses the default process group."""
    import torch
    import torch.distributed as dist
    assert local_slab.dtype == torch.uint8 and local_slab.dim() == 1 and local_slab.numel() == slab_bytes[rank]
    longest = max(slab_bytes) if slab_bytes else 0
    send = torch.zeros(longest, dtype=torch.uint8, device=local_slab.device)
    send[: local_slab.numel()] = local_slab
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    return torch.cat([parts[r][: slab_bytes[r]] for r in range(world)])
