"""Synthetic corpora for the BASELINE.json configs.

pure-zlib has no compressor, so inputs are compressed with the host's system zlib.  All
generators are deterministic (numpy Generator seeds below) and fork-pool parallel; call
them BEFORE the process initialises CUDA.

  config 2  text256k : 4096 x 256 KiB synthetic text, zlib level 6 (dynamic Huffman)
  config 3  records4k: 2^20 x 4 KiB text records, 75 % Z_FIXED / 25 % default strategy
  config 5  stored16m: 512 x 16 MiB random bytes, level 6 (=> ~16 KiB stored blocks)
  config 4  huge     : ONE zlib stream of n MiB of the same text at level 9.  A single-threaded
                       level-9 pass over 1 GiB takes minutes, so the text is compressed in 16 MiB
                       pieces on all cores the way pigz does it: every piece is a raw-deflate run
                       whose compressor was PRIMED with the 32 KiB of text before it (zdict), so its
                       matches reach back across the piece boundary exactly as a one-pass stream's
                       would, and the pieces are joined with Z_SYNC_FLUSH (an empty stored block)
                       behind one zlib header, with one Adler-32 trailer.  The history is never
                       reset: every back-reference of distance <= 32 KiB that a one-pass compressor
                       could have used is available.  (Round 1 ended the pieces with Z_FULL_FLUSH,
                       which cuts the history every 16 MiB: an easier stream than BASELINE states.)
"""
from __future__ import annotations

import hashlib
import multiprocessing as mp
import os
import zlib
from dataclasses import dataclass

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_VOCAB = None


def _vocab():
    """4096 words, length uniform 2..10, letters with English-like (1/rank) frequencies."""
    global _VOCAB
    if _VOCAB is None:
        rng = np.random.default_rng(12345)
        p = 1.0 / np.arange(1, 27)
        p /= p.sum()
        lens = rng.integers(2, 11, 4096)
        flat = rng.choice(_LETTERS, size=int(lens.sum()), p=p)
        off = np.zeros(4097, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        zp = 1.0 / np.arange(1, 4097) ** 1.1
        zp /= zp.sum()
        _VOCAB = (flat, off, lens.astype(np.int64), np.cumsum(zp))
    return _VOCAB


def text(nbytes: int, seed: int) -> bytes:
    """Zipf(1.1) words over the vocabulary; ', ' every 7th word, '.\\n' every 17th, else ' '."""
    flat, off, wlen, cdf = _vocab()
    rng = np.random.default_rng(seed)
    m = nbytes // 5 + 64
    idx = np.searchsorted(cdf, rng.random(m), side="right").clip(0, 4095)
    k = np.arange(1, m + 1)
    sep2 = (k % 17 == 0) | (k % 7 == 0)
    tok = wlen[idx] + 1 + sep2
    ends = np.cumsum(tok)
    starts = ends - tok
    total = int(ends[-1])
    src = np.repeat(off[idx] - starts, tok) + np.arange(total)
    out = np.empty(total, dtype=np.uint8)
    word_mask = np.repeat(wlen[idx], tok) > (np.arange(total) - np.repeat(starts, tok))
    out[word_mask] = flat[src[word_mask]]
    # separators
    s1 = starts + wlen[idx]
    dot = k % 17 == 0
    comma = (k % 7 == 0) & ~dot
    out[s1] = 0x20
    out[s1[dot]] = ord(".")
    out[s1[dot] + 1] = ord("\n")
    out[s1[comma]] = ord(",")
    out[s1[comma] + 1] = 0x20
    assert total >= nbytes
    return out[:nbytes].tobytes()


def _job_text256k(args):
    i, level = args
    d = text(262144, 1000 + i)
    return zlib.compress(d, level), zlib.adler32(d)


def _job_records4k(args):
    lo, hi = args
    out = []
    for i in range(lo, hi):
        d = text(4096, 2_000_000 + i)
        if i % 4 != 3:
            co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
            z = co.compress(d) + co.flush()
        else:
            z = zlib.compress(d, 6)
        out.append((z, zlib.adler32(d)))
    return out


_HUGE_PIECE = 16 << 20


def _job_huge(args):
    k, nbytes, level, last = args
    d = text(nbytes, 3_000_000 + k)
    if k == 0:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
    else:  # the last 32 KiB of the piece before this one are the compressor's history
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, text(_HUGE_PIECE, 3_000_000 + k - 1)[-32768:])
    z = co.compress(d) + (co.flush(zlib.Z_FINISH) if last else co.flush(zlib.Z_SYNC_FLUSH))
    return z, zlib.adler32(d), len(d)


def _job_stored16m(args):
    i, nbytes = args
    d = np.random.default_rng(5_000_000 + i).integers(0, 256, nbytes, dtype=np.uint8).tobytes()
    return zlib.compress(d, 6), zlib.adler32(d)


@dataclass
class Corpus:
    """A batch in the ABI's contiguous layout (offsets 16-byte aligned)."""
    name: str
    in_blob: np.ndarray      # uint8
    in_off: np.ndarray       # uint64, n+1 (in_off[i+1]-in_off[i] includes alignment padding after stream i)
    in_len: np.ndarray       # uint64, n: true compressed length of stream i
    out_len: np.ndarray      # uint64, n: decoded length of stream i
    out_off: np.ndarray      # uint64, n+1
    adler: np.ndarray        # uint32, n
    sha256_in: str

    @property
    def n(self):
        return len(self.out_len)

    @property
    def in_bytes(self):
        return int(self.in_len.sum())

    @property
    def out_bytes(self):
        return int(self.out_len.sum())


def _pack(name, items, out_sizes) -> Corpus:
    n = len(items)
    in_len = np.array([len(z) for z, _ in items], dtype=np.uint64)
    in_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((in_len + np.uint64(15)) & ~np.uint64(15), out=in_off[1:])
    blob = np.zeros(int(in_off[-1]) + 64, dtype=np.uint8)
    h = hashlib.sha256()
    for i, (z, _) in enumerate(items):
        o = int(in_off[i])
        blob[o:o + len(z)] = np.frombuffer(z, dtype=np.uint8)
        h.update(z)
    out_len = np.array(out_sizes, dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((out_len + np.uint64(15)) & ~np.uint64(15), out=out_off[1:])
    adler = np.array([a for _, a in items], dtype=np.uint32)
    return Corpus(name, blob, in_off, in_len, out_len, out_off, adler, h.hexdigest())


def _pool(workers):
    workers = workers or min(os.cpu_count() or 1, 64)
    return mp.get_context("fork").Pool(workers)


def text256k_l1(n: int = 4096, workers: int | None = None) -> "Corpus":
    return text256k(n, 1, workers)


def text256k_l9(n: int = 4096, workers: int | None = None) -> "Corpus":
    return text256k(n, 9, workers)


def text256k(n: int = 4096, level: int = 6, workers: int | None = None) -> Corpus:
    with _pool(workers) as p:
        items = p.map(_job_text256k, [(i, level) for i in range(n)], chunksize=max(1, n // 256))
    return _pack(f"text256k-l{level}", items, [262144] * n)


def records4k(n: int = 1 << 20, workers: int | None = None) -> Corpus:
    step = 2048
    with _pool(workers) as p:
        parts = p.map(_job_records4k, [(lo, min(lo + step, n)) for lo in range(0, n, step)])
    items = [x for part in parts for x in part]
    return _pack("records4k", items, [4096] * n)


def stored16m(n: int = 512, nbytes: int = 16 << 20, workers: int | None = None) -> Corpus:
    with _pool(workers) as p:
        items = p.map(_job_stored16m, [(i, nbytes) for i in range(n)])
    return _pack("stored16m", items, [nbytes] * n)


def huge(n_mib: int = 1024, level: int = 9, workers: int | None = None) -> Corpus:
    """One stream of n_mib MiB (see the module docstring for how it is put together)."""
    total = n_mib << 20
    pieces = [(k, min(_HUGE_PIECE, total - k * _HUGE_PIECE), level, (k + 1) * _HUGE_PIECE >= total)
              for k in range((total + _HUGE_PIECE - 1) // _HUGE_PIECE)]
    with _pool(workers) as p:
        parts = p.map(_job_huge, pieces)
    adler = 1
    for _, a, ln in parts:  # adler32_combine: A = A1 + A2 - 1, B = B1 + B2 + len2 * (A1 - 1)  (mod 65521)
        a1, b1, a2, b2 = adler & 0xffff, adler >> 16, a & 0xffff, a >> 16
        adler = (((b1 + b2 + ln % 65521 * ((a1 + 65520) % 65521)) % 65521) << 16) | ((a1 + a2 + 65520) % 65521)
    z = b"\x78\xda" + b"".join(zz for zz, _, _ in parts) + adler.to_bytes(4, "big")
    return _pack(f"huge{n_mib}m-l{level}", [(z, adler)], [total])


def decoded_piece(k: int, nbytes: int = _HUGE_PIECE) -> bytes:
    """Plain text of piece k of the huge stream."""
    return text(nbytes, 3_000_000 + k)


def decoded_by_name(name: str, i: int, nbytes: int) -> bytes:
    """Regenerates the plain bytes of stream i (its index in the FULL batch) of the corpus called `name`."""
    if name.startswith("text256k"):
        return text(262144, 1000 + i)
    if name == "records4k":
        return text(4096, 2_000_000 + i)
    if name.startswith("huge"):
        return b"".join(decoded_piece(k, min(_HUGE_PIECE, nbytes - k * _HUGE_PIECE)) for k in range((nbytes + _HUGE_PIECE - 1) // _HUGE_PIECE))
    return np.random.default_rng(5_000_000 + i).integers(0, 256, nbytes, dtype=np.uint8).tobytes()


def decoded(corpus: Corpus, i: int) -> bytes:
    """Regenerates the plain text of stream i (for spot checks at full size)."""
    return decoded_by_name(corpus.name, i, int(corpus.out_len[i]))
