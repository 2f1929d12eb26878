"""Builders for hand-made zlib/deflate streams used by the parity tests.

Everything here is test data generation (system zlib as the compressor, plus a bit
writer for streams no compressor would emit).  SURVEY.md Appendix B lists the vectors.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

CODE_LENGTH_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value: int, n: int):
        """n bits, LSB first (RFC 1951 3.1.1 data elements)."""
        for i in range(n):
            self.bits.append((value >> i) & 1)

    def put_code(self, code: int, n: int):
        """Huffman code, MSB first."""
        for i in range(n - 1, -1, -1):
            self.bits.append((code >> i) & 1)

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def put_bytes(self, b: bytes):
        assert len(self.bits) % 8 == 0
        for x in b:
            self.put(x, 8)

    def bytes(self) -> bytes:
        bits = self.bits + [0] * (-len(self.bits) % 8)
        out = bytearray()
        for i in range(0, len(bits), 8):
            v = 0
            for j in range(8):
                v |= bits[i + j] << j
            out.append(v)
        return bytes(out)


def canonical_codes(lens):
    """{sym: (len, code)} per RFC 1951 3.2.2."""
    max_len = max(lens) if len(lens) else 0
    bl = [0] * (max_len + 2)
    for l in lens:
        if l:
            bl[l] += 1
    nc = [0] * (max_len + 2)
    code = 0
    for b in range(1, max_len + 1):
        code = (code + bl[b - 1]) << 1
        nc[b] = code
    out = {}
    for s, l in enumerate(lens):
        if l:
            out[s] = (l, nc[l])
            nc[l] += 1
    return out


FIXED_LIT_LENS = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
FIXED_DIST_LENS = [5] * 32

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
             6145, 8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def zwrap(deflate_body: bytes, data: bytes | None = None, header: bytes = b"\x78\x9c", adler: int | None = None) -> bytes:
    if adler is None:
        adler = zlib.adler32(data if data is not None else b"")
    return header + deflate_body + struct.pack(">I", adler)


class DeflateBuilder:
    """Symbol-level deflate block writer (tokens: int literal, ('m', len, dist), raw symbols)."""

    def __init__(self):
        self.w = BitWriter()

    def _emit_tokens(self, tokens, lit, dist):
        for t in tokens:
            if isinstance(t, int):
                l, c = lit[t]
                self.w.put_code(c, l)
            elif t[0] == "m":
                _, length, d = t
                ls = max(i for i in range(29) if LEN_BASE[i] <= length and (i == 28 or length < LEN_BASE[i] + (1 << LEN_EXTRA[i])))
                if length == 258:
                    ls = 28
                l, c = lit[257 + ls]
                self.w.put_code(c, l)
                self.w.put(length - LEN_BASE[ls], LEN_EXTRA[ls])
                ds = max(i for i in range(30) if DIST_BASE[i] <= d)
                l, c = dist[ds]
                self.w.put_code(c, l)
                self.w.put(d - DIST_BASE[ds], DIST_EXTRA[ds])
            elif t[0] == "litsym":       # raw lit/len symbol (+ optional extra bits)
                l, c = lit[t[1]]
                self.w.put_code(c, l)
                if len(t) > 2:
                    self.w.put(t[2], t[3])
            elif t[0] == "distsym":
                l, c = dist[t[1]]
                self.w.put_code(c, l)
                if len(t) > 2:
                    self.w.put(t[2], t[3])
            elif t[0] == "bits":
                self.w.put(t[1], t[2])
            else:
                raise ValueError(t)

    def fixed(self, tokens, final=True, eob=True):
        self.w.put(1 if final else 0, 1)
        self.w.put(1, 2)
        lit = canonical_codes(FIXED_LIT_LENS)
        dist = canonical_codes(FIXED_DIST_LENS)
        self._emit_tokens(list(tokens) + ([256] if eob else []), lit, dist)
        return self

    def stored(self, data: bytes, final=True, nlen: int | None = None):
        self.w.put(1 if final else 0, 1)
        self.w.put(0, 2)
        self.w.align()
        n = len(data)
        self.w.put(n, 16)
        self.w.put((~n) & 0xFFFF if nlen is None else nlen, 16)
        self.w.put_bytes(data)
        return self

    def dynamic(self, lit_lens, dist_lens, tokens, final=True, eob=True, hlit=None, hdist=None,
                precode_lens=None, raw_length_syms=None):
        """Dynamic block.  By default the code lengths are sent with a flat 4-bit-ish precode
        (symbols 0..15 get length 4: length l is the 4-bit code l), no repeat codes.
        `raw_length_syms` = explicit list of precode symbols (with ('rep', sym, extra, nbits))."""
        w = self.w
        hlit = len(lit_lens) if hlit is None else hlit
        hdist = len(dist_lens) if hdist is None else hdist
        w.put(1 if final else 0, 1)
        w.put(2, 2)
        w.put(hlit - 257, 5)
        w.put(hdist - 1, 5)
        if precode_lens is None:
            precode_lens = [4] * 16 + [0, 0, 0]
        order_lens = [precode_lens[s] for s in CODE_LENGTH_ORDER]
        hclen = 19
        while hclen > 4 and order_lens[hclen - 1] == 0:
            hclen -= 1
        w.put(hclen - 4, 4)
        for i in range(hclen):
            w.put(order_lens[i], 3)
        pc = canonical_codes(precode_lens)
        if raw_length_syms is None:
            raw_length_syms = list(lit_lens) + list(dist_lens)
        for s in raw_length_syms:
            if isinstance(s, int):
                l, c = pc[s]
                w.put_code(c, l)
            else:
                _, sym, extra, nbits = s
                l, c = pc[sym]
                w.put_code(c, l)
                w.put(extra, nbits)
        lit = canonical_codes(list(lit_lens))
        dist = canonical_codes(list(dist_lens))
        self._emit_tokens(list(tokens) + ([256] if eob else []), lit, dist)
        return self

    def body(self) -> bytes:
        return self.w.bytes()


# ----------------------------------------------------------------------------------------
# SURVEY.md Appendix B known-answer vectors: (name, stream bytes, expected)
# expected = ("ok", data) | ("left", show-string) | ("bottom", detail)
# ----------------------------------------------------------------------------------------
def appendix_b_vectors():
    H = bytes.fromhex
    v = []
    hello = b"hello hello hello hello"
    zhello = zlib.compress(hello)
    v.append(("B1", H("789c030000000001"), ("ok", b"")))
    v.append(("B2", H("789c030000000001") + b"JUNK", ("ok", b"")))
    ran_out = "Decompression error: Ran out of data mid-decompression 2."
    for i, s in enumerate([H("789c0300"), H("789c0300000000"), b"", H("78")]):
        v.append((f"B3.{i}", s, ("left", ran_out)))
    v.append(("B4", H("789d030000000001"), ("left", "Header error: Header checksum failed")))
    v.append(("B5", H("7918030000000001"), ("left", "Header error: Bad compression method: 9")))
    v.append(("B6", H("881c030000000001"), ("left", "Header error: Window size too big: 8")))
    v.append(("B7", H("78bbdeadbeef") + zhello[2:], ("ok", hello)))
    v.append(("B8", H("789c0700"), ("left", "Block format error: Unacceptable BTYPE: 3")))
    v.append(("B9", H("789c010500fafe6162636465") + H("05c801f0"),
              ("left", "Block format error: Len/nlen mismatch in uncompressed block.")))
    v.append(("B10", H("789c010500faff6162636465") + H("05c801f0"), ("ok", b"abcde")))
    bad = bytearray(zhello)
    bad[-1] ^= 1
    stored = struct.unpack(">I", bytes(bad[-4:]))[0]
    v.append(("B11", bytes(bad), ("left", "Checksum error: checksum mismatch: %x != %x" % (stored, zlib.adler32(hello)))))
    v.append(("B12", H("789c4b1c030000620062"), ("bottom", 1)))
    v.append(("B13", H("789c4b043e0000000001"), ("bottom", 2)))
    v.append(("B14", H("789c4b04620000000001"), ("bottom", 3)))
    v.append(("B15", H("789c4b1c0500d9a86224"), ("ok", b"a" * 259)))
    v.append(("B16", H("789c4b1cf90000d9a86224"), ("ok", b"a" * 259)))
    # B17: HLIT = 288, HDIST = 32 carrying the fixed code
    b = DeflateBuilder().dynamic(FIXED_LIT_LENS, FIXED_DIST_LENS, [ord("a")])
    v.append(("B17", zwrap(b.body(), b"a"), ("ok", b"a")))
    # B18: incomplete literal code {'a':1 bit, 256: 2 bits}
    ll = [0] * 257
    ll[ord("a")] = 1
    ll[256] = 2
    b = DeflateBuilder().dynamic(ll, [0], [ord("a"), ord("a")])
    v.append(("B18", zwrap(b.body(), b"aa"), ("ok", b"aa")))
    # B19: three 1-bit codes
    ll = [0] * 257
    ll[ord("a")] = 1
    ll[ord("b")] = 1
    ll[256] = 1
    w = DeflateBuilder()
    w.dynamic(ll, [0], [], eob=False)
    v.append(("B19", zwrap(w.body() + b"\0" * 8, b""),
              ("left", "Huffman tree manipulation error: Two values point to the same place!")))
    rng = np.random.default_rng(20)
    rnd = rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes()
    v.append(("B20", zlib.compress(rnd, 0), ("bottom", 4)))
    v.append(("B21a", zlib.compress(rnd, 6), ("ok", rnd)))
    v.append(("B21b", zlib.compress(rnd[:131070], 0), ("ok", rnd[:131070])))
    co = zlib.compressobj(6)
    txt = (b"The quick brown fox jumps over the lazy dog. " * 200)
    s = co.compress(txt[:4000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(txt[4000:]) + co.flush()
    v.append(("B22", s, ("ok", txt)))
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_HUFFMAN_ONLY)
    s = co.compress(rnd[:150_000]) + co.flush()
    v.append(("B23", s, ("ok", rnd[:150_000])))
    co = zlib.compressobj(6, zlib.DEFLATED, 9)
    s = co.compress(txt) + co.flush()
    v.append(("B24", s, ("ok", txt)))
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_DEFAULT_STRATEGY, b"The quick brown fox jumps over the lazy dog. ")
    s = co.compress(txt) + co.flush()
    v.append(("B25", s, ("bottom", 3)))
    return v


# ----------------------------------------------------------------------------------------
# Synthetic text in the style of SURVEY.md 8(d) (small sizes for tests)
# ----------------------------------------------------------------------------------------
def small_text(n: int, seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    vocab = []
    vr = np.random.default_rng(12345)
    p = 1.0 / np.arange(1, 27)
    p /= p.sum()
    for _ in range(512):
        k = int(vr.integers(2, 11))
        vocab.append(bytes(vr.choice(letters, size=k, p=p)))
    out = bytearray()
    ranks = rng.zipf(1.3, size=n // 2 + 16)
    i = 0
    while len(out) < n:
        out += vocab[(int(ranks[i]) - 1) % len(vocab)]
        i += 1
        out += b".\n" if i % 17 == 0 else b", " if i % 7 == 0 else b" "
    return bytes(out[:n])


# ---- gzip / raw deflate (the framing extension): streams and their expected verdict kinds ---------------------------
def gzip_cases():
    """[(name, framing flag name, stream bytes)]: valid members from system zlib (wbits 31 / -15), headers with every
    optional field, and the faults the extension defines verdicts for."""
    import gzip
    import io
    rng = np.random.default_rng(41)
    out = []
    for i, (n, lvl) in enumerate([(0, 6), (1, 9), (5000, 1), (70_000, 6), (300_000, 9)]):
        data = small_text(n, 60 + i)
        co = zlib.compressobj(lvl, zlib.DEFLATED, 31)
        out.append((f"gz-text{n}", "gzip", co.compress(data) + co.flush()))
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15)
        out.append((f"raw-text{n}", "raw", co.compress(data) + co.flush()))
    rnd = rng.integers(0, 256, 90_000, dtype=np.uint8).tobytes()
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    out.append(("gz-stored", "gzip", co.compress(rnd) + co.flush()))
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    out.append(("raw-stored", "raw", co.compress(rnd) + co.flush()))   # ends with the last byte of a stored block
    buf = io.BytesIO()
    with gzip.GzipFile(filename="hello.txt", mode="wb", fileobj=buf, mtime=0) as f:
        f.write(b"hello hello hello")
    g = buf.getvalue()
    out.append(("gz-fname", "gzip", g))
    body = g[10 + len("hello.txt") + 1:]
    hdr = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"EXTRA" + b"name\0" + b"comment\0"
    hdr += (zlib.crc32(hdr) & 0xffff).to_bytes(2, "little")             # FHCRC (the extension skips it, system zlib checks it)
    out.append(("gz-all-fields", "gzip", hdr + body))
    out.append(("gz-two-members", "gzip", g + g))                       # bytes behind the trailer are ignored
    bad = bytearray(g); bad[-5] ^= 1
    out.append(("gz-bad-crc", "gzip", bytes(bad)))
    bad = bytearray(g); bad[-1] ^= 1
    out.append(("gz-bad-isize", "gzip", bytes(bad)))
    out.append(("gz-bad-magic", "gzip", b"\x1f\x8c" + g[2:]))
    out.append(("gz-bad-method", "gzip", g[:2] + b"\x07" + g[3:]))
    out.append(("gz-reserved-flags", "gzip", g[:3] + b"\x80" + g[4:]))
    out.append(("gz-truncated-trailer", "gzip", g[:-3]))
    out.append(("gz-truncated-header", "gzip", g[:12]))
    out.append(("gz-empty", "gzip", b""))
    out.append(("raw-empty", "raw", b""))
    out.append(("raw-truncated", "raw", out[7][2][:-10]))
    out.append(("zlib-as-gzip", "gzip", zlib.compress(b"abc")))
    # one-byte stored blocks, an empty one in between (tools/fuzz_gpu.py found the oracle counting such a byte twice in its CRC);
    # and the same member cut inside its trailer with a wrong CRC in front of the cut: "ran out of data", not a checksum verdict
    import struct
    data = bytes(range(9))
    body = b""
    for i in range(len(data)):
        body += bytes([1 if i == len(data) - 1 else 0]) + struct.pack("<HH", 1, 0xfffe) + data[i:i + 1]
        if i == 3:
            body += b"\x00\x00\x00\xff\xff"
    one = bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 255]) + body + struct.pack("<II", zlib.crc32(data), len(data))
    out.append(("gz-one-byte-stored-blocks", "gzip", one))
    out.append(("gz-bad-crc-cut-in-isize", "gzip", one[:-8] + bytes([one[-8] ^ 1]) + one[-7:-2]))
    return out
