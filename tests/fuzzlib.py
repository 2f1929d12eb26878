"""Corpus + mutation helpers shared by the CPU (host-sim) and GPU differential tests."""
from __future__ import annotations

import zlib

import numpy as np

import streams


def base_corpus(seed: int = 1, count: int = 24):
    """Valid streams spanning stored / fixed / dynamic blocks, several levels and strategies."""
    rng = np.random.default_rng(seed)
    out = []
    for it in range(count):
        n = int(rng.integers(1, 6000))
        kind = it % 4
        if kind == 0:
            data = streams.small_text(n, 1000 + it + seed)
        elif kind == 1:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 2:
            data = (rng.integers(0, 3, n, dtype=np.uint8) * 40 + 65).tobytes()
        else:
            data = bytes(n)
        level = [1, 6, 9, 0][it % 4] if n < 3000 else [1, 6, 9][it % 3]
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED][it % 5]
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        z = co.compress(data[: n // 2])
        if it % 6 == 0:
            z += co.flush(zlib.Z_SYNC_FLUSH)
        if it % 7 == 0:
            z += co.flush(zlib.Z_FULL_FLUSH)
        z += co.compress(data[n // 2:]) + co.flush()
        out.append(z)
    return out


def mutate(z: bytes, rng) -> bytes:
    b = bytearray(z)
    k = int(rng.integers(0, 6))
    if k == 0 and len(b) > 0:      # bit flips
        for _ in range(int(rng.integers(1, 4))):
            i = int(rng.integers(0, len(b)))
            b[i] ^= 1 << int(rng.integers(0, 8))
    elif k == 1:                   # truncate
        b = b[: int(rng.integers(0, len(b) + 1))]
    elif k == 2 and len(b) > 2:    # flip inside the first 40 bytes (headers)
        i = int(rng.integers(2, min(len(b), 40)))
        b[i] ^= 1 << int(rng.integers(0, 8))
    elif k == 3 and len(b) > 0:    # overwrite a byte
        b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif k == 4:                   # append junk
        b += bytes(rng.integers(0, 256, int(rng.integers(1, 8)), dtype=np.uint8))
    else:                          # random garbage after a valid header
        b = bytearray(b"\x78\x9c") + bytes(rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8))
    return bytes(b)


def random_dynamic_stream(rng) -> bytes:
    """A dynamic block with random (often invalid / incomplete / over-subscribed) code lengths
    followed by random bits: exercises tree verdicts and dead prefixes."""
    b = streams.DeflateBuilder()
    w = b.w
    w.put(1, 1)
    w.put(2, 2)
    hlit = int(rng.integers(257, 289))
    hdist = int(rng.integers(1, 33))
    w.put(hlit - 257, 5)
    w.put(hdist - 1, 5)
    mode = int(rng.integers(0, 4))
    if mode == 0:
        pre = [int(x) for x in rng.integers(0, 8, 19)]
    elif mode == 1:
        pre = [4] * 16 + [0, 0, 0]
    elif mode == 2:
        pre = [5] * 13 + [4, 4, 4] + [4, 4, 4]          # complete: 13/32 + 6/16
    else:
        pre = [0] * 19
        for s in rng.choice(19, size=int(rng.integers(1, 6)), replace=False):
            pre[int(s)] = int(rng.integers(1, 4))
    order = [pre[s] for s in streams.CODE_LENGTH_ORDER]
    hclen = int(rng.integers(4, 20))
    w.put(hclen - 4, 4)
    for i in range(hclen):
        w.put(order[i], 3)
    for _ in range(int(rng.integers(0, 400))):
        w.put(int(rng.integers(0, 256)), 8)
    return b"\x78\x9c" + w.bytes()


def structured_dynamic_stream(rng) -> bytes:
    """Dynamic block with a valid flat precode and chosen lit/dist lengths (possibly
    incomplete or over-subscribed), then tokens or noise."""
    nl = int(rng.integers(257, 289))
    nd = int(rng.integers(1, 33))
    style = int(rng.integers(0, 4))
    if style == 0:      # plausible complete-ish code: lengths 7..9
        ll = [int(x) for x in rng.integers(7, 10, nl)]
        dl = [int(x) for x in rng.integers(4, 6, nd)]
    elif style == 1:    # sparse incomplete
        ll = [0] * nl
        for s in rng.choice(nl, size=int(rng.integers(1, 12)), replace=False):
            ll[int(s)] = int(rng.integers(1, 16))
        ll[256] = int(rng.integers(0, 16))
        dl = [0] * nd
        for s in rng.choice(nd, size=int(rng.integers(0, min(nd, 6) + 1)), replace=False):
            dl[int(s)] = int(rng.integers(1, 16))
    elif style == 2:    # long codes (> 10 bits) so the careful walker runs
        ll = [int(x) for x in rng.integers(11, 16, nl)]
        dl = [int(x) for x in rng.integers(9, 16, nd)]
    else:
        ll = [int(x) for x in rng.integers(0, 16, nl)]
        dl = [int(x) for x in rng.integers(0, 16, nd)]
    b = streams.DeflateBuilder()
    toks = []
    lit_ok = [s for s in range(256) if ll[s]]
    len_ok = [s for s in range(257, min(nl, 286)) if ll[s]]
    dist_ok = [s for s in range(min(nd, 30)) if dl[s]]
    produced = 0
    for _ in range(int(rng.integers(0, 300))):
        r = rng.random()
        if lit_ok and (r < 0.6 or not (len_ok and dist_ok and produced)):
            toks.append(int(lit_ok[int(rng.integers(0, len(lit_ok)))]))
            produced += 1
        elif len_ok and dist_ok and produced:
            ls = int(len_ok[int(rng.integers(0, len(len_ok)))])
            ds = int(dist_ok[int(rng.integers(0, len(dist_ok)))])
            le = int(rng.integers(0, 1 << streams.LEN_EXTRA[ls - 257]))
            de = int(rng.integers(0, 1 << streams.DIST_EXTRA[ds]))
            toks.append(("litsym", ls, le, streams.LEN_EXTRA[ls - 257]))
            toks.append(("distsym", ds, de, streams.DIST_EXTRA[ds]))
            produced += streams.LEN_BASE[ls - 257] + le
    try:
        b.dynamic(ll, dl, toks, eob=bool(ll[256]) and rng.random() < 0.8)
    except KeyError:
        return random_dynamic_stream(rng)
    body = b.body()
    tail = bytes(rng.integers(0, 256, int(rng.integers(0, 12)), dtype=np.uint8))
    return b"\x78\x9c" + body + tail


def fuzz_cases(seed: int, n: int):
    rng = np.random.default_rng(seed)
    corpus = base_corpus(seed)
    for i in range(n):
        r = i % 10
        if r < 6:
            yield mutate(corpus[int(rng.integers(0, len(corpus)))], rng)
        elif r < 8:
            yield structured_dynamic_stream(rng)
        else:
            yield random_dynamic_stream(rng)


def big_corpus(seed: int = 1, count: int = 12):
    """Valid streams of 10 KB .. 400 KB: several blocks, long and overlapping matches, the window sliding (the reference
    publishes from 64 KiB on), stored runs between compressed ones."""
    rng = np.random.default_rng(seed)
    out = []
    for it in range(count):
        n = int(rng.integers(10_000, 400_000))
        kind = it % 5
        if kind == 0:
            data = streams.small_text(n, 2000 + it + seed)
        elif kind == 1:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 2:
            data = (rng.integers(0, 4, n, dtype=np.uint8) * 17 + 48).tobytes()
        elif kind == 3:
            unit = rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
            data = (unit * (n // len(unit) + 1))[:n]
        else:
            data = streams.small_text(n // 2, 3000 + it) + rng.integers(0, 256, n - n // 2, dtype=np.uint8).tobytes()
        co = zlib.compressobj(int(rng.integers(1, 10)), zlib.DEFLATED, 15, int(rng.integers(1, 10)),
                              [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED][it % 5])
        z = b""
        cuts = sorted(int(x) for x in rng.integers(0, n, int(rng.integers(0, 4))))
        prev = 0
        for c in cuts:
            z += co.compress(data[prev:c]) + co.flush([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_BLOCK][c % 3])
            prev = c
        z += co.compress(data[prev:]) + co.flush()
        out.append(z)
    return out


def big_fuzz_cases(seed: int, n: int):
    """Mutations of big_corpus streams (every fourth case is the valid stream itself)."""
    rng = np.random.default_rng(seed)
    corpus = big_corpus(seed)
    for i in range(n):
        z = corpus[int(rng.integers(0, len(corpus)))]
        yield z if i % 4 == 0 else mutate(z, rng)


def reframe(z: bytes, framing: int, k: int) -> bytes:
    """The deflate body of a (possibly broken) zlib stream under gzip (1) or raw (2) framing.  gzip members get the trailer of
    what system zlib decodes from the body (a made-up one if it cannot); every third keeps a wrong CRC, every seventh a wrong
    ISIZE, every fifth has an FNAME field."""
    body = z[2:-4] if len(z) >= 6 else z[2:]
    if framing == 2:
        return body
    try:
        data = zlib.decompressobj(-15).decompress(body)
        crc, isz = zlib.crc32(data), len(data) & 0xffffffff
    except Exception:
        crc, isz = 0x12345678, 7
    if k % 3 == 1:
        crc ^= 1 << (k % 32)
    if k % 7 == 3:
        isz ^= 1
    hdr = bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 255])
    if k % 5 == 2:
        hdr = bytes([0x1f, 0x8b, 8, 8, 0, 0, 0, 0, 0, 3]) + b"n\0"
    return hdr + body + crc.to_bytes(4, "little") + isz.to_bytes(4, "little")


def framed_fuzz_cases(seed: int, n: int):
    """(framing, stream): three of five cases zlib, one gzip, one raw deflate (the same generators, re-framed)."""
    for k, z in enumerate(fuzz_cases(seed, n)):
        framing = (0, 0, 0, 1, 2)[k % 5]
        yield framing, (reframe(z, framing, k) if framing else z)


def device_expectation(o):
    """What the inflate kernel alone (before the checksum pass) must report for an oracle
    verdict `o`: the checksum comparison belongs to the Adler kernels."""
    if o.status == 5:
        return (0, 0, 0)
    return (o.status, o.detail, o.payload[0] if o.status in (1, 2, 4, 6) else 0)
