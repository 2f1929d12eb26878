"""CPU suite, part 1: the oracle against the reference's own fixtures and KATs.

test/Test.hs:13-52,107-120 (two canonical-code KATs) and test/Test.hs:56-86 (nine golden
pairs) are the only outputs of the reference that exist; they pin the oracle.  The rest
(Appendix B vectors, zlib cross-checks, chunk semantics) pins the restatement against an
independent derivation (the survey's model / system zlib)."""
import os
import zlib

import numpy as np
import pytest

import streams
from conftest import GOLDEN_NAMES
from oracle import oracle


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden(name, golden_dir):
    z = open(os.path.join(golden_dir, name + ".z"), "rb").read()
    gold = open(os.path.join(golden_dir, name + ".gold"), "rb").read()
    v = oracle.decompress(z)
    assert v.status == 0, v.message
    assert v.data == gold
    assert v.adler_computed == zlib.adler32(gold) == v.adler_stored


def test_kat_rfc1951_code_generation():
    lens = [(ord(c), l) for c, l in zip("ABCDEFGH", [3, 3, 3, 3, 3, 2, 4, 4])]
    want = [(ord("A"), 3, 2), (ord("B"), 3, 3), (ord("C"), 3, 4), (ord("D"), 3, 5), (ord("E"), 3, 6),
            (ord("F"), 2, 0), (ord("G"), 4, 14), (ord("H"), 4, 15)]
    assert oracle.compute_code_values(lens) == want


def test_kat_fixed_huffman():
    lens = [(x, 8) for x in range(144)] + [(x, 9) for x in range(144, 256)] + \
           [(x, 7) for x in range(256, 280)] + [(x, 8) for x in range(280, 288)]
    want = [(x, 8, c) for x, c in zip(range(144), range(48, 192))] + \
           [(x, 9, c) for x, c in zip(range(144, 256), range(400, 512))] + \
           [(x, 7, c) for x, c in zip(range(256, 280), range(0, 24))] + \
           [(x, 8, c) for x, c in zip(range(280, 288), range(192, 200))]
    assert oracle.compute_code_values(lens) == want


@pytest.mark.parametrize("vec", streams.appendix_b_vectors(), ids=lambda v: v[0])
def test_appendix_b(vec):
    name, data, want = vec
    v = oracle.decompress(data)
    if want[0] == "ok":
        assert v.status == 0, (name, v.message)
        assert v.data == want[1]
    elif want[0] == "left":
        assert 1 <= v.status <= 5
        assert v.message == want[1]
    else:
        assert v.status == 6 and v.detail == want[1], (name, v.status, v.detail, v.message)


def test_kraft_rule_matches_trie():
    """accept <=> Kraft sum <= 1; the device relies on it for the verdict class."""
    rng = np.random.default_rng(7)
    seen = {0: 0, 1: 0, 2: 0, 3: 0}
    for it in range(20000):
        n = int(rng.integers(1, 40))
        maxl = int(rng.integers(1, 8))
        lens = rng.integers(0, maxl + 1, n, dtype=np.uint8)
        e, _ = oracle.tree_check(lens.tobytes())
        kraft = sum(2.0 ** -int(l) for l in lens if l)
        assert (e == 0) == (kraft <= 1.0), (lens, e, kraft)
        seen[e] += 1
    assert all(seen[k] > 0 for k in seen)


def test_against_system_zlib_valid_streams():
    rng = np.random.default_rng(11)
    for it in range(60):
        n = int(rng.integers(0, 70000))
        kind = it % 3
        if kind == 0:
            data = streams.small_text(n, 100 + it)
        elif kind == 1:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        else:
            data = (rng.integers(0, 4, n, dtype=np.uint8) * 17).tobytes()
        level = [1, 6, 9][it % 3]
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE][it % 4]
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        z = co.compress(data) + co.flush()
        v = oracle.decompress(z)
        assert v.status == 0 and v.data == data


def test_chunked_input_semantics():
    """Multi-chunk lazy ByteStrings: NeedMore per chunk, 32 KiB Chunk states, trailing chunk error,
    and the getBlock boundary quirk (Monad.hs:279-293)."""
    data = streams.small_text(100_000, 5)
    z = zlib.compress(data, 6)
    whole = oracle.decompress(z, want_events=True)
    assert whole.status == 0
    # split anywhere: same bytes (no stored blocks here)
    for cut in (1, 2, 3, 100, len(z) - 4, len(z) - 1):
        v = oracle.decompress([z[:cut], z[cut:]], want_events=True)
        assert v.status == 0 and v.data == data
        assert [e for e in v.events if e[0] == oracle.EV_NEED_MORE].__len__() == 2
    # chunks of 32 KiB then the remainder
    sizes = [l for k, l in whole.events if k == oracle.EV_CHUNK]
    assert all(s == 32768 for s in sizes[:-1]) and sum(sizes) == len(data)
    # trailing *chunk* after Done
    v = oracle.decompress([z, b"x"])
    assert v.message == "Decompression error: Finished with data remaining."
    # trailing bytes in the same chunk are ignored
    assert oracle.decompress(z + b"xyz").status == 0
    # empty chunks are skipped
    v = oracle.decompress([b"", z[:10], b"", z[10:]])
    assert v.status == 0 and v.data == data
    # quirk: stored block data ending exactly at a chunk boundary swallows one byte
    b = streams.DeflateBuilder().stored(b"abcde", final=True)
    s = streams.zwrap(b.body(), b"abcde")
    end_of_data = 2 + 1 + 4 + 5
    v = oracle.decompress([s[:end_of_data], s[end_of_data:]])
    assert v.status != 0 or v.data != b"abcde"
    v = oracle.decompress([s[:end_of_data + 1], s[end_of_data + 1:]])
    assert v.status == 0 and v.data == b"abcde"


def test_adler():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 5551, 5552, 5553, 70000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.adler32(d) == zlib.adler32(d)
    d = b"\xff" * 200000
    assert oracle.adler32(d) == zlib.adler32(d)


# ---- the framing extension (gzip members, raw deflate): pinned against system zlib on valid streams -----------------
def test_gzip_and_raw_agree_with_system_zlib():
    rng = np.random.default_rng(1)
    for i in range(40):
        n = int(rng.integers(0, 150_000))
        lvl = int(rng.integers(1, 10))
        data = streams.small_text(n, i) if i % 3 else rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        co = zlib.compressobj(lvl, zlib.DEFLATED, 31)
        g = co.compress(data) + co.flush()
        v = oracle.decompress(g, framing=oracle.GZIP)
        assert v.status == 0 and v.data == data and v.adler_computed == zlib.crc32(data) == v.adler_stored, (i, v.message)
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15)
        r = co.compress(data) + co.flush()
        v = oracle.decompress(r, framing=oracle.RAW)
        assert v.status == 0 and v.data == data and v.adler_computed == zlib.adler32(data), (i, v.message)


def test_gzip_verdicts():
    want = {"gz-bad-crc": (5, 1), "gz-bad-isize": (5, 2), "gz-bad-magic": (4, 4), "gz-bad-method": (4, 2), "gz-reserved-flags": (4, 5),
            "gz-truncated-trailer": (3, 1), "gz-truncated-header": (3, 1), "gz-empty": (3, 1), "raw-empty": (3, 1),
            "raw-truncated": (3, 1), "zlib-as-gzip": (4, 4), "gz-bad-crc-cut-in-isize": (3, 1)}
    msgs = {"gz-bad-isize": "Checksum error: length mismatch: 16777233 != 17", "gz-bad-magic": "Header error: Not a gzip stream: 1f8c",
            "gz-reserved-flags": "Header error: Reserved gzip flags set: 128", "gz-bad-method": "Header error: Bad compression method: 7"}
    for name, kind, z in streams.gzip_cases():
        v = oracle.decompress(z, framing=oracle.GZIP if kind == "gzip" else oracle.RAW)
        assert (v.status, v.detail) == want.get(name, (0, 0)), (name, v.status, v.detail, v.message)
        if name in msgs:
            assert v.message == msgs[name]
        if v.status == 0:  # system zlib reads the same bytes (one member / the whole raw stream)
            d = zlib.decompressobj(31 if kind == "gzip" else -15)
            assert d.decompress(z) == v.data
