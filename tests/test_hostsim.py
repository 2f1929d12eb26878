"""CPU suite, part 2: the device decoder LOGIC (pz_device.cuh compiled for the host, one
lane per warp) against the oracle.  This is how kernel logic is iterated without a GPU; the
CUDA build itself is checked by the `-m gpu` tests."""
import os
import zlib

import pytest

import fuzzlib
import streams
from conftest import GOLDEN_NAMES
from hostsim import hostsim  # noqa: E402
from oracle import oracle


def check(data: bytes):
    o = oracle.decompress(data)
    cap = o.out_len + 300
    r, out = hostsim.inflate(data, cap)
    want = fuzzlib.device_expectation(o)
    got = (r.status, r.detail, r.payload[0] if r.status in (1, 2, 4, 6) else 0)
    assert got == want, (data.hex()[:200], o.message, got, want)
    assert r.out_len == o.out_len, (o.message, r.out_len, o.out_len)
    assert out == o.data[: len(out)]
    if o.status in (0, 5):
        assert r.adler_stored == o.adler_stored
    if o.status == 6 and o.detail == 3:
        assert r.payload[1] == o.payload[1]
    # the sizing pass reaches the same verdict and length without writing
    r2, _ = hostsim.inflate(data, cap, count_only=True)
    assert (r2.status, r2.detail, r2.out_len) == (r.status, r.detail, r.out_len)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden(name, golden_dir):
    check(open(os.path.join(golden_dir, name + ".z"), "rb").read())


@pytest.mark.parametrize("vec", streams.appendix_b_vectors(), ids=lambda v: v[0])
def test_appendix_b(vec):
    check(vec[1])


def test_valid_corpus():
    for z in fuzzlib.base_corpus(3, 40):
        check(z)


@pytest.mark.parametrize("seed", range(8))
def test_fuzz(seed):
    for data in fuzzlib.fuzz_cases(seed, 400):
        check(data)


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_big_streams(seed):
    """Mutations of 10 KB .. 400 KB streams (several blocks, long and overlapping matches, the window sliding)."""
    for data in fuzzlib.big_fuzz_cases(seed, 24):
        check(data)


def test_output_full():
    data = streams.small_text(5000, 1)
    z = zlib.compress(data)
    r, out = hostsim.inflate(z, 4999)
    assert r.status == 7
    r, out = hostsim.inflate(z, 5000)
    assert r.status == 0 and out == data


def test_smem_budget():
    # one CTA per SM holds 28 stream slots: at most 227 KiB of dynamic shared memory per CTA, and
    # the slot stride must be 16 (mod 128) bytes so the hot warp's lanes hit different banks
    n = hostsim.lib().hs_smem_bytes()
    assert n * 28 <= 227 * 1024 and n % 128 == 16


def _resume_cuts(z: bytes, rng, k: int):
    cuts = sorted(set(int(x) for x in rng.integers(0, len(z) + 1, k)) | {len(z)})
    return cuts


def check_resumed(z: bytes, cuts):
    """PzJob::resume / PzJob::ckpt (the incremental driver's device contexts): decoding prefix after
    prefix, each run picking up at the previous run's checkpoint, must give exactly what a decode of
    the whole prefix from its first byte gives -- verdict, length, published bytes, output."""
    full = oracle.decompress(z)
    ctx = hostsim.Resumable(full.out_len + 300)
    kinds = set()
    for cut in cuts:
        prefix = z[:cut]
        o = oracle.decompress(prefix)
        r, out, used = ctx.run(prefix)
        kinds.add("fresh" if used[0] == 0 else "trailer" if used[1] == 0xFFFFFFFF else "header" if used[1] == 0 else "symbol")
        want = fuzzlib.device_expectation(o)
        got = (r.status, r.detail, r.payload[0] if r.status in (1, 2, 4, 6) else 0)
        assert got == want, (cut, used, o.message, got, want)
        assert r.out_len == o.out_len, (cut, used, r.out_len, o.out_len)
        assert out == o.data[: len(out)], (cut, used)
        if o.status == 3 and o.detail == 1:  # NeedMore: the chunks the reference has handed out by now
            assert r.payload[1] == o.published, (cut, used, r.payload[1], o.published)
        if o.status in (0, 5):
            assert r.adler_stored == o.adler_stored
        if not (o.status == 3 and o.detail == 1):  # anything but "ran out of data" is final
            break
    return kinds


@pytest.mark.parametrize("seed", range(4))
def test_resume_small(seed):
    import numpy as np
    rng = np.random.default_rng(77 + seed)
    for z in fuzzlib.base_corpus(10 + seed, 30):
        check_resumed(z, _resume_cuts(z, rng, 6))
        check_resumed(z, list(range(0, min(len(z), 64))) + [len(z)])  # every byte of the first headers


def test_resume_multiblock():
    """Streams long enough for the window to slide (published bytes move) and with several blocks:
    checkpoints inside blocks, at headers, and in the trailer."""
    import numpy as np
    rng = np.random.default_rng(5)
    data = streams.small_text(300_000, 9)
    for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY)):
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        z = co.compress(data[:100_000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(data[100_000:]) + co.flush()
        kinds = check_resumed(z, _resume_cuts(z, rng, 12))
        assert "symbol" in kinds
        kinds = check_resumed(z, [len(z) - 5, len(z) - 4, len(z) - 3, len(z) - 1, len(z)])  # the trailer arrives byte by byte
        assert "trailer" in kinds
    # stored blocks only (level 6 on random bytes: ~16 KiB stored blocks)
    z = zlib.compress(np.random.default_rng(6).integers(0, 256, 200_000, dtype=np.uint8).tobytes(), 6)
    check_resumed(z, _resume_cuts(z, rng, 12))


@pytest.mark.parametrize("seed", range(4))
def test_resume_fuzz(seed):
    """Mutated streams: whatever verdict the whole prefix has, the resumed decode has it too."""
    import numpy as np
    rng = np.random.default_rng(900 + seed)
    for data in fuzzlib.fuzz_cases(seed, 150):
        if len(data) < 2:
            continue
        check_resumed(data, _resume_cuts(data, rng, 4))


def test_gzip_and_raw_framing():
    """The framing extension in the device logic (header parse, trailer words, the stored-block rule of raw deflate)
    against the oracle's: everything but the checksum comparison, which is K3's on the device."""
    for name, kind, z in streams.gzip_cases():
        fr = 1 if kind == "gzip" else 2
        o = oracle.decompress(z, framing=fr)
        for count_only in (False, True):
            r, out = hostsim.inflate(z, o.out_len + 300, count_only, framing=fr)
            if o.status == 5:  # the decoder reads the trailer; comparing is the checksum kernels' business
                assert (r.status, r.out_len, r.adler_stored) == (0, o.out_len, o.adler_stored), name
                continue
            assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (name, count_only, r.status, r.detail, r.out_len, o.message)
            if not count_only:
                assert out == o.data, name
            if o.status == 0 and fr == 1:
                assert r.adler_stored == o.adler_stored and r.payload[0] == o.out_len, name


@pytest.mark.parametrize("seed", [21, 22])
def test_gzip_and_raw_framing_fuzz(seed):
    """Mutated and generated deflate bodies under gzip and raw framing (fuzzlib.reframe: right and wrong CRC / ISIZE, FNAME
    fields, truncated trailers): the device logic against the oracle's extension -- found by tools/fuzz_gpu.py: a member cut
    inside its trailer is "ran out of data" whatever its CRC says."""
    n = 0
    for fr, z in fuzzlib.framed_fuzz_cases(seed, 1500):
        if fr == 0:
            continue
        o = oracle.decompress(z, framing=fr)
        r, out = hostsim.inflate(z, o.out_len + 300, False, framing=fr)
        n += 1
        if o.status == 5:  # comparing checksums is K3's business on the device
            assert (r.status, r.out_len, r.adler_stored) == (0, o.out_len, o.adler_stored), (seed, n, z.hex()[:80])
            continue
        assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (seed, n, r.status, r.detail, r.out_len, o.message, z.hex()[:80])
        assert out == o.data
    assert n >= 500


# ---- K5 (pz_fixed.cuh): one thread per small fixed-Huffman stream -----------------------------------------------------
def _fixed_stream(data: bytes, level: int = 6, blocks: int = 1) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    step = max(1, (len(data) + blocks - 1) // blocks)
    out = b""
    for k in range(0, max(len(data), 1), step):
        out += co.compress(data[k:k + step])
        if k + step < len(data):
            out += co.flush(zlib.Z_BLOCK)     # ends the block, no empty stored block: the next one is fixed again
    return out + co.flush()


def _check_fixed(z: bytes, expect_taken=None):
    """What K5 completes must be the oracle's success, field by field; what the oracle does not call a success
    K5 must leave to K1 (it never reports anything but success)."""
    o = oracle.decompress(z, want_events=True)
    for count_only, out_mis in ((False, 0), (False, 1 + len(z) % 3), (True, 0)):
        # (exact capacity when the stream is fine: the last word of the output is then an incomplete one)
        ok, r, out = hostsim.fixed(z, o.out_len if o.status == 0 else max(o.out_len, 1) + 64, count_only, out_mis)
        if o.status != 0:
            assert not ok, (z.hex()[:120], o.message)
            continue
        if expect_taken is not None:
            assert ok == expect_taken, (expect_taken, z.hex()[:80])
        if not ok:
            continue
        published = sum(ln for kind, ln in o.events[:-2] if kind == 1)
        assert (r.status, r.detail, r.out_len, r.adler_stored, r.payload[1]) == (0, 0, o.out_len, o.adler_stored, published), \
            (r.out_len, o.out_len, r.payload[1], published)
        if not count_only:
            assert out == o.data
        # err_bitpos is what K1 reports for the same stream
        r1, _ = hostsim.inflate(z, o.out_len + 64)
        assert r.err_bitpos == r1.err_bitpos
    return o


def test_fixed_kernel_logic_valid_streams():
    import numpy as np
    rng = np.random.default_rng(3)
    taken = 0
    for i in range(120):
        n = int(rng.integers(0, 9000))
        kind = i % 4
        data = streams.small_text(n, i) if kind < 2 else rng.integers(0, 256, n, dtype=np.uint8).tobytes() if kind == 2 else bytes(n)
        z = _fixed_stream(data, int(rng.integers(1, 10)), int(rng.integers(1, 4)))
        if len(z) <= 16384:
            # (random bytes leave zlib as STORED blocks even with Z_FIXED: K5 rightly leaves those alone)
            _check_fixed(z, expect_taken=None if kind == 2 else True)
            taken += kind != 2
    assert taken > 80
    # FDICT is skipped (Zlib.hs:68): the header's check bits have to fit the flag
    z = bytearray(_fixed_stream(b"hello hello hello"))
    body = bytes(z[2:])
    hdr = bytes([0x78, 0xbb])
    assert (hdr[0] * 256 + hdr[1]) % 31 == 0
    o = _check_fixed(hdr + b"\xde\xad\xbe\xef" + body, expect_taken=True)
    assert o.data == b"hello hello hello"
    # streams K5 must not take although they are fine: a dynamic or stored first block, a dynamic block further on, too long
    _check_fixed(zlib.compress(streams.small_text(3000, 1), 6), expect_taken=False)
    _check_fixed(zlib.compress(rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(), 6), expect_taken=False)
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    z = co.compress(streams.small_text(2000, 2)) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(b"tail") + co.flush()
    _check_fixed(z, expect_taken=False)  # the flush leaves an empty stored block between the fixed ones
    _check_fixed(_fixed_stream(streams.small_text(60000, 9)), expect_taken=False)
    # more than 64 KiB of output (the window slides, payload[1] moves) does not fit 16 KiB of fixed-coded text; zeros do
    _check_fixed(_fixed_stream(bytes(200_000)), expect_taken=True)


@pytest.mark.parametrize("seed", range(4))
def test_fixed_kernel_logic_fuzz(seed):
    """Mutated fixed-Huffman streams: whatever K5 completes is the oracle's success; every other verdict is K1's."""
    import numpy as np
    rng = np.random.default_rng(100 + seed)
    done = left = 0
    for i in range(300):
        n = int(rng.integers(1, 3000))
        data = streams.small_text(n, 1000 * seed + i) if i % 3 else bytes(rng.integers(0, 4, n, dtype=np.uint8))
        z = bytearray(_fixed_stream(data, 6, int(rng.integers(1, 3))))
        how = i % 5
        if how == 0:
            z = z[: int(rng.integers(0, len(z)))]                       # truncated
        elif how == 1:
            for _ in range(int(rng.integers(1, 4))):
                z[int(rng.integers(0, len(z)))] ^= 1 << int(rng.integers(0, 8))   # bit flips
        elif how == 2:
            z[-1] ^= 0xff                                                # checksum (K3's verdict: K5 completes the stream)
        elif how == 3:
            z += bytes(rng.integers(0, 256, int(rng.integers(0, 9)), dtype=np.uint8))  # bytes behind the trailer
        z = bytes(z)
        o = oracle.decompress(z, want_events=True)
        ok, r, out = hostsim.fixed(z, max(o.out_len, 1) + 64, False, i % 4)
        if o.status in (0, 5):  # success, or success up to the checksum comparison
            if ok:
                assert (r.out_len, r.adler_stored) == (o.out_len, o.adler_stored) and out == o.data
                done += 1
        else:
            assert not ok, (z.hex()[:100], o.message, r.out_len)
            left += 1
    assert done > 60 and left > 40, (done, left)


# ---- K6: the same one-thread-per-stream decoder with per-thread tables for dynamic blocks --------------------------------
def _check_small(z: bytes):
    """K6 (pz_fixed_stream<.., DYN = true>): what it completes is the oracle's success; anything else it leaves to K1."""
    o = oracle.decompress(z, want_events=True)
    took = False
    for count_only, out_mis in ((False, 0), (False, 1 + len(z) % 3), (True, 0)):
        ok, r, out = hostsim.fixed(z, o.out_len if o.status == 0 else max(o.out_len, 1) + 64, count_only, out_mis, dyn=True)
        if o.status not in (0, 5):
            assert not ok, (z.hex()[:120], o.message)
            continue
        if not ok:
            continue
        took = True
        assert (r.out_len, r.adler_stored) == (o.out_len, o.adler_stored)
        if o.status == 0:
            assert r.payload[1] == sum(ln for kind, ln in o.events[:-2] if kind == 1)
            r1, _ = hostsim.inflate(z, o.out_len + 64)
            assert r.err_bitpos == r1.err_bitpos
        if not count_only:
            assert out == o.data
    return o, took


def test_small_stream_dynamic_logic():
    import numpy as np
    from pure_zlib_b200 import corpus
    rng = np.random.default_rng(9)
    taken = total = 0
    for i in range(150):
        n = int(rng.integers(1, 12000))
        kind = i % 5
        data = streams.small_text(n, 500 + i) if kind < 2 else corpus.text(n, 7_000_000 + i) if kind < 4 else bytes(rng.integers(0, 7, n, dtype=np.uint8))
        co = zlib.compressobj(int(rng.integers(1, 10)), zlib.DEFLATED, 15, int(rng.integers(1, 10)),
                              [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE][i % 4])
        z = co.compress(data[: n // 3]) + (co.flush(zlib.Z_BLOCK) if i % 2 else b"") + co.compress(data[n // 3:]) + co.flush()
        if len(z) > 16384:
            continue
        o, took = _check_small(z)
        assert o.status == 0
        total += 1
        taken += took
    assert taken >= 0.8 * total, (taken, total)   # what it leaves: blocks with codes longer than its tables' index
    # the benchmark's records: how many of the dynamic ones does K6 take?
    got = 0
    for i in range(3, 400, 4):
        z = zlib.compress(corpus.text(4096, 2_000_000 + i), 6)
        got += _check_small(z)[1]
    assert got >= 90, got


@pytest.mark.parametrize("seed", range(4))
def test_small_stream_dynamic_fuzz(seed):
    import numpy as np
    rng = np.random.default_rng(300 + seed)
    done = left = 0
    cases = list(fuzzlib.fuzz_cases(40 + seed, 250))
    for i in range(150):
        n = int(rng.integers(1, 5000))
        z = bytearray(zlib.compress(streams.small_text(n, 9000 + 200 * seed + i), int(rng.integers(1, 10))))
        how = i % 4
        if how == 0:
            z = z[: int(rng.integers(0, len(z)))]
        elif how == 1:
            for _ in range(int(rng.integers(1, 4))):
                z[int(rng.integers(0, len(z)))] ^= 1 << int(rng.integers(0, 8))
        elif how == 2:
            z[-2] ^= 0x10
        cases.append(bytes(z))
    for z in cases:
        if len(z) > 16384:
            continue
        o, took = _check_small(z)
        done += took
        left += o.status not in (0, 5)
    assert done > 40 and left > 40, (done, left)
