"""GPU suite: the CUDA path, called through the C ABI, against the oracle (bit-exact bytes
and verdicts).  Mirrors test/Test.hs (2 KATs + 9 goldens) and widens to the behaviours the
reference's own tests do not reach."""
import ctypes as C
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

import fuzzlib
import streams
from conftest import GOLDEN_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pz():
    import pure_zlib_b200 as pz
    return pz


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


def verdict_of(pz, data):
    try:
        return pz.decompress(data)
    except pz.ReferenceBottom as e:
        return ("bottom", str(e))


def expect_of(pz, o):
    if o.status == 0:
        return pz.Right(o.data)
    if o.status == 6:
        return ("bottom", o.message)
    return ("left", o.message)


def same(pz, got, o):
    if o.status == 0:
        return got == pz.Right(o.data)
    if o.status == 6:
        return got == ("bottom", o.message)
    return isinstance(got, pz.Left) and str(got.value) == o.message


# ---- test/Test.hs -----------------------------------------------------------------------
def test_kat_rfc1951_code_generation(pz):
    lens = [(ord(c), l) for c, l in zip("ABCDEFGH", [3, 3, 3, 3, 3, 2, 4, 4])]
    want = [(ord("A"), 3, 2), (ord("B"), 3, 3), (ord("C"), 3, 4), (ord("D"), 3, 5), (ord("E"), 3, 6),
            (ord("F"), 2, 0), (ord("G"), 4, 14), (ord("H"), 4, 15)]
    assert pz.compute_code_values(lens) == want


def test_kat_fixed_huffman(pz):
    lens = [(x, 8) for x in range(144)] + [(x, 9) for x in range(144, 256)] + \
           [(x, 7) for x in range(256, 280)] + [(x, 8) for x in range(280, 288)]
    want = [(x, 8, c) for x, c in zip(range(144), range(48, 192))] + \
           [(x, 9, c) for x, c in zip(range(144, 256), range(400, 512))] + \
           [(x, 7, c) for x, c in zip(range(256, 280), range(0, 24))] + \
           [(x, 8, c) for x, c in zip(range(280, 288), range(192, 200))]
    assert pz.compute_code_values(lens) == want


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden(pz, name, golden_dir):
    z = open(os.path.join(golden_dir, name + ".z"), "rb").read()
    gold = open(os.path.join(golden_dir, name + ".gold"), "rb").read()
    assert pz.decompress(z) == pz.Right(gold)


# ---- beyond the reference's tests --------------------------------------------------------
@pytest.mark.parametrize("vec", streams.appendix_b_vectors(), ids=lambda v: v[0])
def test_appendix_b(pz, oracle, vec):
    name, data, want = vec
    o = oracle.decompress(data)
    got = verdict_of(pz, data)
    assert same(pz, got, o), (name, got, o.message)
    if want[0] == "left":
        assert str(got.value) == want[1]


def test_batch_matches_oracle_mixed_verdicts(pz, oracle):
    """One launch holding valid, malformed and truncated streams: each verdict is the
    oracle's, and a bad stream does not disturb its neighbours."""
    cases = []
    for seed in range(6):
        cases += list(fuzzlib.fuzz_cases(seed, 500))
    cases += fuzzlib.base_corpus(9, 60)
    res, outs = pz.zlib.inflate_batch_raw(cases)
    from pure_zlib_b200 import _lib
    bad = []
    for i, (data, r, out) in enumerate(zip(cases, res, outs)):
        o = oracle.decompress(data)
        ok = r.status == o.status and r.detail == o.detail and r.out_len == o.out_len and out == o.data
        if o.status in (1, 2, 4, 6):
            ok = ok and r.payload[0] == o.payload[0]
        if o.status in (0, 5):
            ok = ok and r.adler_computed == o.adler_computed and r.adler_stored == o.adler_stored
        if o.status != 0:
            ok = ok and _lib.strerror(r) == o.message
        if not ok:
            bad.append((i, data.hex()[:80], o.message, (r.status, r.detail, r.out_len), (o.status, o.detail, o.out_len)))
    assert not bad, bad[:5]


def test_valid_streams_levels_and_strategies(pz):
    rng = np.random.default_rng(5)
    cases, plain = [], []
    for it in range(64):
        n = int(rng.integers(0, 300_000))
        data = streams.small_text(n, it) if it % 2 == 0 else (rng.integers(0, 7, n, dtype=np.uint8) * 31).tobytes()
        level = [1, 6, 9][it % 3]
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE][it % 4]
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        cases.append(co.compress(data) + co.flush())
        plain.append(data)
    got = pz.decompress_batch(cases)
    for g, d in zip(got, plain):
        assert g == pz.Right(d)


def test_empty_and_ragged_batch(pz):
    a = zlib.compress(b"")
    b = zlib.compress(b"x" * 100_000)
    got = pz.decompress_batch([a, b, b"", a, b[:-3]])
    assert got[0] == pz.Right(b"") and got[3] == pz.Right(b"")
    assert got[1] == pz.Right(b"x" * 100_000)
    assert str(got[2].value) == "Decompression error: Ran out of data mid-decompression 2."
    assert str(got[4].value) == "Decompression error: Ran out of data mid-decompression 2."
    assert pz.decompress_batch([]) == []


def test_large_expansion_and_output_full(pz):
    from pure_zlib_b200 import _lib
    L = _lib.load()
    data = bytes(3_000_000)
    z = zlib.compress(data, 9)
    assert pz.decompress(z) == pz.Right(data)
    # capacity one byte short -> PZ_OUTPUT_FULL, never a wrong verdict
    inb = C.create_string_buffer(z, len(z))
    out = C.create_string_buffer(len(data))
    ptr = (C.c_void_p * 1)(C.addressof(inb))
    ln = (C.c_size_t * 1)(len(z))
    optr = (C.c_void_p * 1)(C.addressof(out))
    cap = (C.c_size_t * 1)(len(data) - 1)
    res = (_lib.PzResult * 1)()
    _lib.check(L.pz_inflate_batch(ptr, ln, optr, cap, 1, res, 0), "pz_inflate_batch")
    assert res[0].status == _lib.PZ_OUTPUT_FULL


def test_adler32_entry_point(pz):
    from pure_zlib_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(2)
    for n in (0, 1, 15, 16, 17, 5551, 5552, 16384, 16385, 1_000_003):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert L.pz_adler32(1, d, n) == zlib.adler32(d)
    d = b"\xff" * 70000
    assert L.pz_adler32(1, d, len(d)) == zlib.adler32(d)
    assert L.pz_adler32(zlib.adler32(b"abc"), b"defgh", 5) == zlib.adler32(b"abcdefgh")


def test_incremental_event_sequence(pz, oracle):
    """decompressIncremental: the NeedMore / Chunk / Done sequence is the reference's."""
    data = streams.small_text(200_000, 77)
    z = zlib.compress(data, 6)
    for cuts in ([len(z)], [1, 2, len(z) - 3], [5000, 9000, 20000], [len(z) - 4], [len(z) - 1]):
        pieces, prev = [], 0
        for c in cuts:
            pieces.append(z[prev:c])
            prev = c
        if prev < len(z):
            pieces.append(z[prev:])
        o = oracle.decompress(pieces, want_events=True)
        events = []
        st = pz.decompress_incremental()
        rest = list(pieces)
        acc = b""
        while True:
            if isinstance(st, pz.NeedMore):
                events.append((0, 0))
                if not rest:
                    break
                st = st.feed(rest.pop(0))
            elif isinstance(st, pz.Chunk):
                events.append((1, len(st.data)))
                acc += st.data
                st = st.next()
            elif isinstance(st, pz.Done):
                events.append((2, 0))
                break
            else:
                events.append((3, 0))
                break
        assert events == o.events, (cuts, events[:8], o.events[:8])
        assert acc == data


def test_multichunk_decompress(pz):
    data = streams.small_text(50_000, 3)
    z = zlib.compress(data, 6)
    assert pz.decompress([z[:100], z[100:]]) == pz.Right(data)
    got = pz.decompress([z, b"tail"])
    assert str(got.value) == "Decompression error: Finished with data remaining."
    got = pz.decompress([z[:100], z[100:200]])
    assert str(got.value) == "Decompression error: Ran out of data mid-decompression 2."


def test_contig_device_pointers_and_resident_batch(pz):
    """The launch-only path bench.py times: device-resident blobs, pz_batch_*."""
    import torch
    from pure_zlib_b200 import _lib, corpus
    L = _lib.load()
    c = corpus.text256k(48, workers=4)
    d_in = torch.from_numpy(c.in_blob).cuda()
    d_out = torch.zeros(int(c.out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    in_off = np.ascontiguousarray(c.in_off.copy())
    # true stream ends (not padded) are what the decoder may read
    ends = c.in_off[:-1] + c.in_len
    res = (_lib.PzResult * c.n)()
    b = L.pz_batch_create(in_off.ctypes.data_as(C.POINTER(C.c_uint64)), c.out_off.ctypes.data_as(C.POINTER(C.c_uint64)), c.n, 0)
    assert b
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.pz_batch_run(b, d_in.data_ptr(), d_out.data_ptr(), st), "pz_batch_run")
    _lib.check(L.pz_batch_results(b, res, st), "pz_batch_results")
    assert L.pz_batch_launches(b) == 5  # K2 probe, K2 copy, K1, K3a, K3b
    L.pz_batch_destroy(b)
    out = d_out.cpu().numpy()
    for i in range(c.n):
        assert res[i].status == 0, (i, res[i].status, res[i].detail)
        assert res[i].out_len == 262144 and res[i].adler_computed == c.adler[i]
        o = int(c.out_off[i])
        assert out[o:o + 262144].tobytes() == corpus.decoded(c, i)
    assert ends is not None


# ---- K2: streams of stored blocks only ------------------------------------------------------
def _stored_cases():
    """Incompressible data (level 6 => chains of ~16 KiB stored blocks) and every way such a
    stream can stop being the fast path's business."""
    rng = np.random.default_rng(11)
    rnd = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()  # noqa: E731
    cases = []
    for n in (0, 1, 5, 4096, 16383, 16384, 65535, 65536, 131070, 200_000, 1_500_000):
        cases.append(zlib.compress(rnd(n), 6))
    big = zlib.compress(rnd(300_000), 6)
    cases.append(zlib.compress(rnd(200_000), 0))            # 65535-byte blocks: window overflow (bottom) at the third
    cases.append(zlib.compress(rnd(131_070), 0))            # two 65535-byte blocks: fine
    co = zlib.compressobj(6)
    cases.append(co.compress(rnd(50_000)) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(rnd(50_000)) + co.flush())  # empty stored blocks
    co = zlib.compressobj(6)
    cases.append(co.compress(rnd(40_000)) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(streams.small_text(40_000, 5)) + co.flush())  # stored, then Huffman
    cases.append(big[:-1]); cases.append(big[:-4]); cases.append(big[:-5]); cases.append(big[:50_000])  # truncations
    b = bytearray(big); b[-1] ^= 1; cases.append(bytes(b))  # checksum mismatch
    b = bytearray(big); b[2 + 3] ^= 0x10; cases.append(bytes(b))  # NLEN of the first block
    # LEN of the third block (walk the chain to find it)
    p, k = 2, 0
    while k < 2:
        ln = big[p + 1] | (big[p + 2] << 8)
        p += 5 + ln; k += 1
    b = bytearray(big); b[p + 1] ^= 1; cases.append(bytes(b))
    b = bytearray(big); b[p] |= 0x06; cases.append(bytes(b))  # BTYPE 3 in the third block
    co = zlib.compressobj(0)  # many small stored blocks (and the empty ones of the flushes): several pieces per checksum segment
    cases.append(b"".join(co.compress(rnd(n)) + co.flush(zlib.Z_SYNC_FLUSH) for n in [1, 2, 3, 700, 1000, 15, 16, 17, 5000] * 12) + co.flush())
    cases.append(big + b"trailing junk")
    cases.append(bytes([0x78, 0xbb]) + b"\xde\xad\xbe\xef" + big[2:])  # FDICT set: four bytes skipped
    return cases


def test_stored_streams_fast_path_and_fallbacks(pz, oracle):
    from pure_zlib_b200 import _lib
    cases = _stored_cases()
    res, outs = pz.zlib.inflate_batch_raw(cases)
    for i, (data, r, out) in enumerate(zip(cases, res, outs)):
        o = oracle.decompress(data)
        assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (i, len(data), o.message, r.status, r.detail, r.out_len)
        assert out == o.data, i
        if o.status in (0, 5):
            assert (r.adler_computed, r.adler_stored) == (o.adler_computed, o.adler_stored), i
        if o.status != 0:
            assert _lib.strerror(r) == o.message, i


def test_stored_output_capacity(pz):
    from pure_zlib_b200 import _lib
    L = _lib.load()
    data = np.random.default_rng(3).integers(0, 256, 100_000, dtype=np.uint8).tobytes()
    z = zlib.compress(data, 6)
    for cap, want in ((len(data), _lib.PZ_OK), (len(data) - 1, _lib.PZ_OUTPUT_FULL), (20_000, _lib.PZ_OUTPUT_FULL)):
        inb = C.create_string_buffer(z, len(z))
        out = C.create_string_buffer(len(data))
        res = (_lib.PzResult * 1)()
        _lib.check(L.pz_inflate_batch((C.c_void_p * 1)(C.addressof(inb)), (C.c_size_t * 1)(len(z)), (C.c_void_p * 1)(C.addressof(out)),
                                      (C.c_size_t * 1)(cap), 1, res, 0), "pz_inflate_batch")
        assert res[0].status == want, (cap, res[0].status)
        if want == _lib.PZ_OK:
            assert out.raw == data


# ---- the BASELINE configs at scale: size-independent properties --------------------------------
def _run_corpus(c):
    """Decodes a corpus through the resident-batch path; returns (verdict records, device output)."""
    import torch
    from pure_zlib_b200 import _lib
    L = _lib.load()
    p64 = C.POINTER(C.c_uint64)
    d_in = torch.from_numpy(c.in_blob).cuda()
    d_out = torch.zeros(int(c.out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    b = L.pz_batch_create(c.in_off.ctypes.data_as(p64), c.out_off.ctypes.data_as(p64), c.n, 0)
    assert b
    st = torch.cuda.current_stream().cuda_stream
    res = (_lib.PzResult * c.n)()
    _lib.check(L.pz_batch_run(b, d_in.data_ptr(), d_out.data_ptr(), st), "pz_batch_run")
    _lib.check(L.pz_batch_results(b, res, st), "pz_batch_results")
    L.pz_batch_destroy(b)
    rec = np.frombuffer(res, dtype=np.dtype([("status", "<i4"), ("detail", "<i4"), ("out_len", "<u8"), ("adler_c", "<u4"),
                                             ("adler_s", "<u4"), ("bitpos", "<u8"), ("p0", "<i8"), ("p1", "<i8")]))
    return rec, d_out


@pytest.mark.parametrize("name,n", [("text256k", 512), ("records4k", 200_000), ("stored16m", 12)])
def test_baseline_config_properties(pz, name, n):
    """Every stream OK, decoded length and Adler-32 equal to the generator's (the checksum of each
    output is the domain's size-independent check), spot streams byte-for-byte."""
    from pure_zlib_b200 import corpus
    c = getattr(corpus, name)(n, workers=8)
    rec, d_out = _run_corpus(c)
    assert (rec["status"] == 0).all(), np.unique(rec["status"], return_counts=True)
    assert (rec["out_len"] == c.out_len).all()
    assert (rec["adler_c"] == c.adler).all() and (rec["adler_s"] == c.adler).all()
    assert (rec["bitpos"] == c.in_len * 8).all()
    host = d_out.cpu().numpy()
    for i in np.linspace(0, c.n - 1, 8).astype(int):
        o = int(c.out_off[i])
        assert host[o:o + int(c.out_len[i])].tobytes() == corpus.decoded(c, int(i)), (name, i)
    if name == "records4k":  # a sample against the oracle: fixed-Huffman records (K5) and dynamic ones (K1) alike
        from oracle import oracle as orc
        for i in np.linspace(0, c.n - 1, 300).astype(int):
            z = bytes(c.in_blob[int(c.in_off[i]): int(c.in_off[i]) + int(c.in_len[i])])
            v = orc.decompress(z, want_events=True)
            o = int(c.out_off[i])
            assert v.status == 0 and host[o:o + int(c.out_len[i])].tobytes() == v.data, i
            assert (int(rec["out_len"][i]), int(rec["adler_c"][i]), int(rec["p1"][i])) == \
                (v.out_len, v.adler_computed, sum(ln for kind, ln in v.events[:-2] if kind == 1)), i


def test_small_stream_batch_mixed_verdicts(pz, oracle):
    """A batch big enough for K5 (one thread per small fixed-Huffman stream, pz_fixed.cuh) whose streams are of every
    kind: fixed, dynamic and stored, several blocks, and faults (truncated, flipped bits, bad checksum, symbols the
    reference cannot index).  K5 only ever reports success; everything else must come out of K1 with the reference's
    verdict -- through the decode, the one-call decode and the sizing pass."""
    from pure_zlib_b200 import _lib
    rng = np.random.default_rng(77)
    base = []
    for i in range(300):
        n = int(rng.integers(0, 6000))
        data = streams.small_text(n, 3000 + i) if i % 4 else bytes(rng.integers(0, 3, n, dtype=np.uint8))
        strat = [zlib.Z_FIXED, zlib.Z_FIXED, zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY][i % 4]
        co = zlib.compressobj(int(rng.integers(1, 10)), zlib.DEFLATED, 15, 8, strat)
        z = co.compress(data[: n // 2]) + (co.flush(zlib.Z_BLOCK) if i % 3 == 0 else b"") + co.compress(data[n // 2:]) + co.flush()
        base.append(z)
    faults = []
    for i, z in enumerate(base[:120]):
        b = bytearray(z)
        if i % 4 == 0 and len(b) > 8:
            b = b[: int(rng.integers(1, len(b)))]
        elif i % 4 == 1 and len(b) > 8:
            b[int(rng.integers(2, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif i % 4 == 2:
            b[-1] ^= 0x55
        else:
            b += b"trailing"
        faults.append(bytes(b))
    faults += [v[1] for v in streams.appendix_b_vectors()]
    cases = (base + faults) * 24           # 10 000+ streams: above PZ_FIXED_MIN_STREAMS
    assert len(cases) >= 8192
    want = {}
    for fn in (pz.zlib.decompress_batch_raw, pz.zlib.inflate_batch_raw):
        res, outs = fn(cases)
        for i in list(range(len(base) + len(faults))) + [len(cases) - 1 - k for k in range(50)]:
            z = cases[i]
            o = want.setdefault(z, oracle.decompress(z))
            r = res[i]
            assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (i, r.status, r.detail, r.out_len, o.message)
            assert outs[i] == o.data, i
            if o.status in (0, 5):
                assert (r.adler_computed, r.adler_stored) == (o.adler_computed, o.adler_stored), i
            if o.status != 0:
                assert _lib.strerror(r) == o.message, i


def test_small_stream_batch_with_k6_opt_in():
    """K6 (the thread-per-stream kernel with private dynamic-code tables) is opt-in (PZ_K6=1, read once per process): the same
    mixed batch in a child process that has it on."""
    if os.environ.get("PZ_K6"):
        pytest.skip("already the child")
    env = dict(os.environ, PZ_K6="1")
    env.pop("PZ_NO_K6", None)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k", "small_stream_batch_mixed_verdicts",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]


def test_sizing_pass_matches_decode(pz):
    """pz_inflate_sizes (decoder warps only) reaches the same length and verdict as the decode."""
    from pure_zlib_b200 import _lib
    L = _lib.load()
    cases = fuzzlib.base_corpus(4, 40) + _stored_cases()[:8] + list(fuzzlib.fuzz_cases(3, 200))
    res, _ = pz.zlib.inflate_batch_raw(cases)
    n = len(cases)
    bufs = [C.create_string_buffer(z, max(len(z), 1)) for z in cases]
    ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
    lens = (C.c_size_t * n)(*[len(z) for z in cases])
    sz = (_lib.PzResult * n)()
    _lib.check(L.pz_inflate_sizes(ptrs, lens, n, sz), "pz_inflate_sizes")
    for i in range(n):
        want_status = 0 if res[i].status == 5 else res[i].status  # the sizing pass stops before the checksum compare
        assert (sz[i].status, sz[i].out_len) == (want_status, res[i].out_len), (i, sz[i].status, res[i].status)


# ---- K4: one huge stream decoded block-parallel -------------------------------------------------
@pytest.fixture()
def huge_threshold():
    """Lowers the size from which a stream takes the block-parallel path, so that streams the
    oracle finishes in seconds exercise it; restores the default afterwards."""
    from pure_zlib_b200 import _lib
    L = _lib.load()
    _lib.check(L.pz_init(None), "pz_init")
    assert L.pz_set_option(_lib.PZ_OPT_HUGE_BYTES, 64 * 1024) == 0
    yield
    assert L.pz_set_option(_lib.PZ_OPT_HUGE_BYTES, 4 << 20) == 0


def _huge_cases():
    from pure_zlib_b200 import corpus
    text = corpus.text(6 << 20, 77)
    cases = [("text-l9", zlib.compress(text, 9)), ("text-l6", zlib.compress(text, 6)), ("text-l1", zlib.compress(text[: 2 << 20], 1))]
    # a fixed-Huffman block and a stored block in the middle: blocks the header search cannot see
    co = zlib.compressobj(9)
    z = co.compress(text[: 1 << 20]) + co.flush(zlib.Z_FULL_FLUSH)
    z += co.compress(text[1 << 20: 2 << 20]) + co.flush()
    cases.append(("sync-flush", z))
    # a stream compressed in pieces: the empty stored blocks of Z_SYNC_FLUSH are candidates of the search too
    co = zlib.compressobj(9)
    cases.append(("many-flushes", b"".join(co.compress(text[k << 19: (k + 1) << 19]) + co.flush(zlib.Z_SYNC_FLUSH) for k in range(8)) + co.flush()))
    cf = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    cases.append(("all-fixed", cf.compress(text[: 1 << 20]) + cf.flush()))
    rng = np.random.default_rng(5)
    mixed = text[: 1 << 20] + rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes() + text[1 << 20: 2 << 20]
    cases.append(("mixed-random", zlib.compress(mixed, 9)))
    # faults: truncated, corrupted in the middle, bad checksum, trailing garbage
    good = cases[0][1]
    cases.append(("truncated", good[: len(good) // 2]))
    bad = bytearray(good); bad[len(bad) // 3] ^= 0x40
    cases.append(("corrupt", bytes(bad)))
    bad = bytearray(good); bad[-1] ^= 1
    cases.append(("checksum", bytes(bad)))
    cases.append(("trailing", good + b"xyz"))
    # references before the start of the stream: a valid tail glued behind a fresh header
    return cases


def test_huge_stream_block_parallel(pz, oracle, huge_threshold):
    """Streams above the threshold take K4 (or fall back): bytes, verdict and the published-bytes
    model must equal the oracle's, and equal the serial path's (PZ_F_NO_HUGE)."""
    from pure_zlib_b200 import _lib
    cases = _huge_cases()
    L = _lib.load()
    done0, declined0 = L.pz_get_counter(1), L.pz_get_counter(2)
    res, outs = pz.zlib.inflate_batch_raw([z for _, z in cases])
    # the seven well-formed streams, the one with a bad checksum and the one with bytes after its trailer take K4
    # (inflate_batch_raw decodes twice: sizing pass without K4, then the decode); the broken ones are declined
    # (a flipped bit may leave every block decodable -- then K4 finishes and the checksum fails -- or not)
    done, declined = L.pz_get_counter(1) - done0, L.pz_get_counter(2) - declined0
    assert done + declined == len(cases) and done >= 9 and declined >= 1, (done, declined)
    res0, outs0 = pz.zlib.inflate_batch_raw([z for _, z in cases], flags=_lib.PZ_F_NO_HUGE)
    for (name, z), r, out, r0, out0 in zip(cases, res, outs, res0, outs0):
        o = oracle.decompress(z)
        assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (name, r.status, r.detail, o.message)
        assert (r0.status, r0.detail, r0.out_len) == (o.status, o.detail, o.out_len), name
        assert out == o.data and out0 == o.data, name
        if o.status == 0:
            assert r.adler_computed == o.adler_computed == r0.adler_computed, name
            assert r.payload[1] == r0.payload[1], (name, r.payload[1], r0.payload[1])  # published-bytes model (serial path = hostsim-checked)
            assert r.err_bitpos == r0.err_bitpos, name
        else:
            assert _lib.strerror(r) == o.message, name


def test_huge_stream_cross_piece_references_64mib(pz, oracle):
    """BASELINE configs[3] in small: ONE 64 MiB level-9 stream whose 16 MiB pieces reference each other across
    their boundaries (corpus.huge: primed compressors joined with Z_SYNC_FLUSH, the history is never reset), at
    the library's DEFAULT threshold (4 MiB compressed): K4 must take it, and every piece must be byte-exact
    against the generator and the whole against the oracle."""
    from pure_zlib_b200 import _lib, corpus
    L = _lib.load()
    c = corpus.huge(64)
    z = bytes(c.in_blob[: int(c.in_len[0])])
    assert len(z) >= 4 << 20
    done0 = L.pz_get_counter(1)
    n = 1
    keep = C.create_string_buffer(z, len(z))
    ptrs = (C.c_void_p * n)(C.addressof(keep))
    lens = (C.c_size_t * n)(len(z))
    cap = int(c.out_len[0])
    out = C.create_string_buffer(cap)
    optrs = (C.c_void_p * n)(C.addressof(out))
    caps = (C.c_size_t * n)(cap)
    res = (_lib.PzResult * n)()
    _lib.check(L.pz_inflate_batch(ptrs, lens, optrs, caps, n, res, 0), "pz_inflate_batch")
    assert L.pz_get_counter(1) - done0 == 1, "K4 declined the stream"
    o = oracle.decompress(z, out_cap=cap, want_events=True)
    assert o.status == 0 and (res[0].status, res[0].out_len, res[0].adler_computed) == (0, o.out_len, o.adler_computed)
    assert res[0].adler_computed == int(c.adler[0])
    got = out.raw
    piece = 16 << 20
    for k in range(cap // piece):
        assert got[k * piece: (k + 1) * piece] == corpus.decoded_piece(k), f"piece {k} differs"
    assert got == o.data
    assert res[0].payload[1] == sum(ln for kind, ln in o.events[:-2] if kind == 1)
    # the serial path agrees (verdict fields the block-parallel path reconstructs)
    res0 = (_lib.PzResult * n)()
    _lib.check(L.pz_inflate_batch(ptrs, lens, optrs, caps, n, res0, _lib.PZ_F_NO_HUGE), "pz_inflate_batch")
    assert (res0[0].status, res0[0].out_len, res0[0].adler_computed, res0[0].err_bitpos, res0[0].payload[1]) == \
        (0, res[0].out_len, res[0].adler_computed, res[0].err_bitpos, res[0].payload[1])
    assert out.raw == o.data


def test_huge_stream_window_model_with_generous_capacity(pz, oracle, huge_threshold):
    """The reference's 128 KiB window overflows on back-to-back 64 KiB stored blocks wherever they
    sit in a stream.  Block jobs do not know the window's fill at their block's start, so K4 must give such a
    stream to the serial path (PzCtx::mark: a gap of more than 32 KiB between two moveWindow calls) -- also when
    the caller's capacities would let it finish.  Published-bytes counts (payload[1]) must agree too."""
    from pure_zlib_b200 import _lib, corpus
    L = _lib.load()
    rng = np.random.default_rng(9)
    full = corpus.text((1 << 20) + 20000, 78)
    raw = rng.integers(0, 256, 65535, dtype=np.uint8).tobytes()

    def stored(data, final):
        return bytes([1 if final else 0]) + len(data).to_bytes(2, "little") + (len(data) ^ 0xffff).to_bytes(2, "little") + data

    def stream(parts, text=full):
        co = zlib.compressobj(9, zlib.DEFLATED, -15)
        body = co.compress(text) + co.flush(zlib.Z_SYNC_FLUSH)  # ends on a byte boundary, not final
        plain = text
        for k, d in enumerate(parts):
            body += stored(d, k == len(parts) - 1)
            plain += d
        return b"\x78\xda" + body + zlib.adler32(plain).to_bytes(4, "big"), plain

    # whether the second 65535-byte block overflows depends on the window's fill when the first one starts, i.e. on
    # the length of everything before it modulo 32 KiB: 1 MiB + 20000 bytes overflow, 1 MiB exactly does not
    cases = [stream([raw, raw]), stream([raw, raw], full[: 1 << 20]), stream([raw[:40000], raw[:20000]]),
             stream([raw[:32768], raw[:32768], raw[:100]]), stream([raw[:32769], raw[:5]])]
    n = len(cases)
    keep = [C.create_string_buffer(z, len(z)) for z, _ in cases]
    ptrs = (C.c_void_p * n)(*[C.addressof(k) for k in keep])
    lens = (C.c_size_t * n)(*[len(z) for z, _ in cases])
    caps = (C.c_size_t * n)(*[len(p) + 4096 for _, p in cases])
    outs = [C.create_string_buffer(len(p) + 4096) for _, p in cases]
    optrs = (C.c_void_p * n)(*[C.addressof(o) for o in outs])
    for flags in (0, _lib.PZ_F_NO_HUGE):
        res = (_lib.PzResult * n)()
        _lib.check(L.pz_inflate_batch(ptrs, lens, optrs, caps, n, res, flags), "pz_inflate_batch")
        for i, (z, plain) in enumerate(cases):
            o = oracle.decompress(z, want_events=True)
            r = res[i]
            assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (flags, i, r.status, r.detail, r.out_len, o.message)
            assert outs[i].raw[: r.out_len] == o.data, (flags, i)
            if o.status == 0:
                published = sum(ln for kind, ln in o.events[:-2] if kind == 1)  # every Chunk but the final one
                assert r.payload[1] == published, (flags, i, r.payload[1], published)
    assert oracle.decompress(cases[0][0]).status == 6  # the first case is the advisor's: window overflow


def test_huge_stream_in_a_batch_and_device_pointers(pz, huge_threshold):
    """A huge stream among small ones, through the contiguous entry point with device pointers."""
    import torch
    from pure_zlib_b200 import _lib, corpus
    L = _lib.load()
    datas = [corpus.text(3 << 20, 5), b"small one" * 10, corpus.text(1 << 20, 6), b""]
    zs = [zlib.compress(d, 9) for d in datas]
    n = len(zs)
    in_off = np.zeros(n + 1, dtype=np.uint64); out_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([(len(z) + 15) & ~15 for z in zs], out=in_off[1:])
    np.cumsum([(len(d) + 15) & ~15 for d in datas], out=out_off[1:])
    blob = np.zeros(int(in_off[-1]) + 64, dtype=np.uint8)
    for i, z in enumerate(zs):
        blob[int(in_off[i]): int(in_off[i]) + len(z)] = np.frombuffer(z, dtype=np.uint8)
    # in_off[i+1]-in_off[i] includes padding: the padding bytes are "data after the trailer", ignored like the reference does
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(int(out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    res = (_lib.PzResult * n)()
    p64 = C.POINTER(C.c_uint64)
    _lib.check(L.pz_inflate_batch_contig(d_in.data_ptr(), in_off.ctypes.data_as(p64), d_out.data_ptr(), out_off.ctypes.data_as(p64), n, res, None, 0),
               "pz_inflate_batch_contig")
    host = d_out.cpu().numpy()
    for i, d in enumerate(datas):
        assert res[i].status == 0 and res[i].out_len == len(d) and res[i].adler_computed == zlib.adler32(d), i
        assert host[int(out_off[i]): int(out_off[i]) + len(d)].tobytes() == d, i


# ---- resumable device contexts (pz_stream_*: checkpoints, one launch for many streams) -------
def _pieces(z, cuts):
    out, prev = [], 0
    for c in sorted(set(cuts)):
        if c > prev:
            out.append(z[prev:c])
            prev = c
    if prev < len(z):
        out.append(z[prev:])
    return out


def _run_incremental(pz, pieces):
    events, acc, st, rest, err = [], b"", pz.decompress_incremental(), list(pieces), None
    while True:
        if isinstance(st, pz.NeedMore):
            events.append((0, 0))
            if not rest:  # `decompress`'s driver loop: "Ran out of data mid-decompression 2." (Zlib.hs:38-39)
                events.append((3, 0))
                break
            st = st.feed(rest.pop(0))
        elif isinstance(st, pz.Chunk):
            events.append((1, len(st.data)))
            acc += st.data
            st = st.next()
        elif isinstance(st, pz.Done):
            events.append((2, 0))  # with chunks left over the driver says "Finished with data remaining." (Zlib.hs:48-49)
            break
        else:
            events.append((3, 0))
            err = st.error
            break
    return events, acc, err


def test_incremental_resumes_from_checkpoint(pz, oracle):
    """A stream fed in many pieces is decoded once, not once per piece: every pump starts at the
    checkpoint of the one before (inside a block, at a header, in the trailer), and the events are
    still the reference's."""
    from pure_zlib_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(11)
    data = streams.small_text(400_000, 5)
    co = zlib.compressobj(6)
    z = co.compress(data[:150_000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(data[150_000:]) + co.flush()
    stored = zlib.compress(rng.integers(0, 256, 150_000, dtype=np.uint8).tobytes(), 6)
    fixed = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    zf = fixed.compress(data[:90_000]) + fixed.flush()
    for stream in (z, stored, zf):
        # (no cut at len - 4 for the stored stream: a chunk boundary exactly at the end of a stored block's data
        # is the reference's getBlock quirk, SURVEY Appendix A.5, outside the engine's contract)
        tail = [len(stream) - 5, len(stream) - 3, len(stream) - 2, len(stream) - 1] if stream is stored else \
            [len(stream) - 5, len(stream) - 4, len(stream) - 2, len(stream) - 1]
        for cuts in (list(range(0, len(stream), 4099)), [int(x) for x in rng.integers(0, len(stream), 9)], tail):
            pieces = _pieces(stream, cuts)
            o = oracle.decompress(pieces, want_events=True)
            events, acc, err = _run_incremental(pz, pieces)
            assert events == o.events, (len(stream), cuts[:6], events[:8], o.events[:8])
            assert acc == o.data[: len(acc)] and (o.status != 0 or acc == o.data)
    # the decoder really resumes: decoded bytes behind the checkpoint grow with the input
    d = pz.zlib._Decoder()
    seen = []
    for piece in _pieces(z, list(range(0, len(z), 10_007))):
        _lib.check(L.pz_stream_feed(d._s, piece, len(piece)), "feed")
        st = d.state()
        while isinstance(st, pz.Chunk):
            st = st.next()
        seen.append((L.pz_stream_counter(d._s, _lib.PZ_SC_CKPT_BIT), L.pz_stream_counter(d._s, _lib.PZ_SC_CKPT_BYTES)))
    assert isinstance(st, pz.Done)
    pumps, resumed = L.pz_stream_counter(d._s, _lib.PZ_SC_PUMPS), L.pz_stream_counter(d._s, _lib.PZ_SC_RESUMED)
    assert resumed >= pumps - 1 >= len(seen) - 2, (pumps, resumed)
    bits = [b for b, _ in seen[:-1]]
    assert bits == sorted(bits) and bits[-1] > 8 * (len(z) - 2 * 10_007 - 300), bits[-3:]
    # each checkpoint is at most one symbol (<= 48 bits) plus what the next piece starts with behind the input
    for k, (b, _) in enumerate(seen[:-1]):
        assert 8 * min(len(z), 10_007 * (k + 1)) - b <= 64, (k, b)


def test_incremental_verdicts_fuzz(pz, oracle):
    """Mutated streams in random pieces: the terminal state (and the bytes handed out before it) are
    what the reference's `decompress` over the same chunk list gives."""
    rng = np.random.default_rng(21)
    n = 0
    for data in fuzzlib.fuzz_cases(5, 60):
        if len(data) < 4:
            continue
        pieces = _pieces(data, [int(x) for x in rng.integers(1, len(data), 3)])
        o = oracle.decompress(pieces, want_events=True)
        if o.status == 6:
            with pytest.raises(pz.ReferenceBottom):
                _run_incremental(pz, pieces)
            continue
        events, acc, err = _run_incremental(pz, pieces)
        assert events == o.events, (data.hex()[:80], events[:6], o.events[:6], o.message)
        if o.status not in (0, 3) or (o.status == 3 and o.detail == 2):
            assert err is not None and str(err) == o.message
        n += 1
    assert n > 20


def test_stream_pump_many(pz, oracle):
    """pz_stream_pump: 96 concurrent incremental consumers, one launch per round of chunks."""
    from pure_zlib_b200 import _lib
    rng = np.random.default_rng(31)
    files, want = [], []
    for i in range(96):
        kind = i % 4
        if kind == 0:
            raw = streams.small_text(int(rng.integers(1, 200_000)), 100 + i)
        elif kind == 1:
            raw = rng.integers(0, 256, int(rng.integers(1, 90_000)), dtype=np.uint8).tobytes()
        elif kind == 2:
            raw = bytes(int(rng.integers(1, 300_000)))
        else:
            raw = streams.small_text(int(rng.integers(1, 5000)), 900 + i)
        z = zlib.compress(raw, [1, 6, 9][i % 3])
        if i % 11 == 5:   # a corrupted trailer
            z = z[:-1] + bytes([z[-1] ^ 1])
        if i % 13 == 7:   # truncated: the driver loop runs out of chunks
            z = z[: len(z) // 2]
        k = int(rng.integers(1, 7))
        pieces = _pieces(z, [int(x) for x in rng.integers(1, max(2, len(z)), k)])
        if i % 17 == 3:
            pieces.append(b"junk")  # data after the end of the stream
        files.append(pieces)
        want.append(oracle.decompress(pieces))
    got = pz.decompress_many(files)
    for i, (g, o) in enumerate(zip(got, want)):
        if o.status == 6:
            assert isinstance(g, pz.ReferenceBottom) and str(g) == o.message, i
        else:
            assert same(pz, g, o), (i, g, o.message)
    # the set API: every stream decoded by exactly one launch per round it was fed in
    group = pz.IncrementalSet(8)
    zs = [zlib.compress(streams.small_text(120_000, 40 + i), 6) for i in range(8)]
    rounds = 5
    outs = [b""] * 8
    for r in range(rounds):
        parts = [z[r * ((len(z) + rounds - 1) // rounds):(r + 1) * ((len(z) + rounds - 1) // rounds)] for z in zs]
        if r % 2 == 0:  # one call for the round (pz_stream_feed_many) ...
            group.feed_all(parts)
        else:           # ... or one per stream
            for i, part in enumerate(parts):
                group.feed(i, part)
        group.pump()
        for i in range(8):
            for st in group.events(i):
                if isinstance(st, pz.Chunk):
                    outs[i] += st.data
    for i in range(8):
        assert isinstance(st, pz.Done)
        assert outs[i] == streams.small_text(120_000, 40 + i)
        assert group.counter(i, _lib.PZ_SC_PUMPS) == rounds
        assert group.counter(i, _lib.PZ_SC_RESUMED) == rounds - 1


def test_incremental_feed_sizes(pz, oracle):
    """Feeds far above the staging block (1 MiB: the chunk travels in pieces), and a stream fed one byte at a time
    (every header field, code length and symbol is cut somewhere)."""
    rng = np.random.default_rng(41)
    big = zlib.compress(rng.integers(0, 256, 3_500_000, dtype=np.uint8).tobytes(), 6)      # stored blocks, 3.5 MB in one chunk
    text = zlib.compress(streams.small_text(6_000_000, 8), 6)                                # ~1.9 MB compressed
    for z, cuts in ((big, []), (big, [2_000_001]), (text, [1_500_000]), (text, [])):
        pieces = _pieces(z, cuts)
        o = oracle.decompress(pieces, want_events=True)
        events, acc, err = _run_incremental(pz, pieces)
        assert events == o.events and acc == o.data and err is None, (len(z), cuts, events[:4], o.events[:4])
    small = zlib.compress(streams.small_text(3000, 12), 9)
    fixed = zlib.compressobj(9, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    smallf = fixed.compress(streams.small_text(1200, 13)) + fixed.flush()
    for z in (small, smallf):
        pieces = [z[i:i + 1] for i in range(len(z))]
        o = oracle.decompress(pieces, want_events=True)
        events, acc, err = _run_incremental(pz, pieces)
        assert events == o.events and acc == o.data and err is None, (len(z), events[-4:], o.events[-4:])


def test_stream_feed_many_arguments(pz):
    from pure_zlib_b200 import _lib
    L = _lib.load()
    group = pz.IncrementalSet(3)
    with pytest.raises(_lib.PzCudaError):   # a stream twice in one call
        group.feed_all([b"a", b"b"], [1, 1])
    group.feed_all([], [])                  # nothing to do
    z = zlib.compress(b"abc" * 1000)
    group.feed_all([z, b"", z[:10]])        # an empty chunk is accepted and ignored
    group.pump()
    kinds = [[type(st).__name__ for st in group.events(i)] for i in range(3)]
    assert kinds[0] == ["Chunk", "Done"] and kinds[1] == ["NeedMore"] and kinds[2] == ["NeedMore"], kinds


# ---- thread safety (pzcuda.h: every entry point may be called concurrently; SURVEY 8(b): `decompress` is pure) ----
def test_concurrent_host_threads(pz, oracle):
    """Eight host threads hammer the batch entry point, single incremental decoders and pumped sets at the same
    time (ctypes releases the GIL for the duration of every call, so the calls really overlap inside the
    library: thread-local workspaces, the process-wide pinned cache, the shared CUDA-stream pool).  Every verdict
    and every byte must be the oracle's, whichever thread produced it."""
    import threading
    from pure_zlib_b200 import _lib
    rng = np.random.default_rng(31)
    texts = [streams.small_text(int(rng.integers(1000, 120_000)), 100 + i) for i in range(24)]
    cases = [zlib.compress(t, int(rng.integers(1, 10))) for t in texts]
    cases += [zlib.compress(rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes(), 6)]      # stored blocks (K2)
    cases += [cases[0][:-3], bytes.fromhex("789c4b04620000000001"), b"", cases[1][:100] + b"\xff" * 40]  # verdicts other than OK
    want = [oracle.decompress(z, want_events=True) for z in cases]
    errors = []

    def check_batch(tid, rounds):
        for r in range(rounds):
            order = list(np.random.default_rng(1000 * tid + r).permutation(len(cases)))
            res, outs = pz.zlib.inflate_batch_raw([cases[i] for i in order])
            for k, i in enumerate(order):
                o = want[i]
                if (res[k].status, res[k].detail, res[k].out_len) != (o.status, o.detail, o.out_len) or outs[k] != o.data:
                    errors.append(("batch", tid, r, i, res[k].status, res[k].detail, o.status, o.detail))
                elif o.status != 0 and _lib.strerror(res[k]) != o.message:
                    errors.append(("batch-msg", tid, r, i, _lib.strerror(res[k]), o.message))

    def check_incremental(tid, rounds):
        for r in range(rounds):
            i = (tid * 7 + r) % len(cases)
            if want[i].status == 6 or len(cases[i]) < 8:
                continue
            cuts = sorted(set(int(x) for x in np.random.default_rng(tid * 77 + r).integers(1, len(cases[i]), 3)))
            pieces = _pieces(cases[i], cuts)
            o = oracle.decompress(pieces, want_events=True)
            events, acc, err = _run_incremental(pz, pieces)
            if events != o.events or acc != o.data[: len(acc)]:
                errors.append(("incremental", tid, r, i, events[:4], o.events[:4]))

    def check_pump(tid, rounds):
        ok = [i for i in range(len(cases)) if want[i].status == 0 and len(cases[i]) > 64]
        for r in range(rounds):
            pick = [ok[(tid + r + 3 * k) % len(ok)] for k in range(6)]
            got = pz.zlib.decompress_many([[cases[i][: len(cases[i]) // 2], cases[i][len(cases[i]) // 2:]] for i in pick])
            for i, g in zip(pick, got):
                if not isinstance(g, pz.Right) or g.value != want[i].data:
                    errors.append(("pump", tid, r, i, repr(g)[:60]))

    workers = []
    for tid in range(8):
        fn = (check_batch, check_incremental, check_pump)[tid % 3]
        workers.append(threading.Thread(target=fn, args=(tid, 6 if fn is check_batch else 8)))
    for w in workers:
        w.start()
    for w in workers:
        w.join(timeout=600)
    assert not any(w.is_alive() for w in workers), "a worker thread hangs"
    assert not errors, errors[:5]


def test_deflate_cli(pz, golden_dir, tmp_path):
    """The `deflate` executable (reference Deflate.hs:15-48; haskell/app/Deflate.hs; pure_zlib_b200/deflate_cli.py):
    `deflate foo.z` writes foo through the incremental API, with the reference's messages."""
    import shutil
    from pure_zlib_b200 import deflate_cli
    said = []
    for name in ("rfctest2", "zerotest3", "randtest1"):
        shutil.copy(os.path.join(golden_dir, name + ".z"), tmp_path / (name + ".z"))
        assert deflate_cli.main([str(tmp_path / (name + ".z"))], said.append) == 0
        assert (tmp_path / name).read_bytes() == open(os.path.join(golden_dir, name + ".gold"), "rb").read()
    assert said == []
    z = open(os.path.join(golden_dir, "rfctest1.z"), "rb").read()
    (tmp_path / "cut.z").write_bytes(z[:5000])
    deflate_cli.main([str(tmp_path / "cut.z")], said.append)
    (tmp_path / "bad.z").write_bytes(bytes.fromhex("789c0700"))
    deflate_cli.main([str(tmp_path / "bad.z")], said.append)
    (tmp_path / "more.z").write_bytes(z + b"\0" * 40000)   # a further lazy chunk after Done
    deflate_cli.main([str(tmp_path / "more.z")], said.append)
    deflate_cli.main([str(tmp_path / "name.txt")], said.append)
    deflate_cli.main([], said.append)
    assert said == ["ERROR: Ran out of data mid-decompression.", "ERROR: Block format error: Unacceptable BTYPE: 3",
                    "WARNING: Finished decompression with data left.", "Unexpected file name.", "USAGE: deflate [filename]"]
    assert (tmp_path / "more").read_bytes() == open(os.path.join(golden_dir, "rfctest1.gold"), "rb").read()


def test_decompress_batch_single_call(pz, oracle):
    """pz_decompress_batch (sizing, allocation and decode inside the library: what the Haskell shim's decompressBatch
    binds) against the two-call flow and the oracle, verdicts other than OK included."""
    from pure_zlib_b200 import _lib
    cases = [v[1] for v in streams.appendix_b_vectors()] + [zlib.compress(streams.small_text(70_000, 3), 6), b"",
                                                            zlib.compress(b"", 9), zlib.compress(streams.small_text(300_000, 4), 1)[:-1]]
    r1, o1 = pz.zlib.decompress_batch_raw(cases)
    r2, o2 = pz.zlib.inflate_batch_raw(cases)
    for i, z in enumerate(cases):
        o = oracle.decompress(z)
        assert (r1[i].status, r1[i].detail, r1[i].out_len) == (o.status, o.detail, o.out_len) == (r2[i].status, r2[i].detail, r2[i].out_len), i
        assert o1[i] == o.data == o2[i], i
        if o.status != 0:
            assert _lib.strerror(r1[i]) == o.message


def test_in_library_multi_gpu(pz):
    """pz_config.devices / PZ_DEVICES: ONE call from ONE host thread shards a host batch over every device (contiguous
    ranges balanced by compressed bytes, a worker thread per device).  Needs two GPUs; the library initialises once
    per process, so the sharded run happens in a child process."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    code = r'''
import sys, zlib, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import streams
import pure_zlib_b200 as pz
from pure_zlib_b200 import _lib
from oracle import oracle
L = _lib.load()
cfg = _lib.PzConfig(); cfg.device = -1; cfg.n_devices = -1
assert L.pz_init(__import__("ctypes").byref(cfg)) == 0 and L.pz_device_count() >= 2, L.pz_device_count()
rng = np.random.default_rng(3)
cases = [zlib.compress(streams.small_text(int(rng.integers(100, 200_000)), i), int(rng.integers(1, 10))) for i in range(97)]
cases[5] = cases[5][:-2]; cases[40] = bytes.fromhex("789c4b04620000000001"); cases[96] = b""
cases.append(zlib.compress(rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes(), 6))
for fn in (pz.zlib.decompress_batch_raw, pz.zlib.inflate_batch_raw):
    res, outs = fn(cases)
    for i, z in enumerate(cases):
        o = oracle.decompress(z)
        assert (res[i].status, res[i].detail, res[i].out_len) == (o.status, o.detail, o.out_len), (i, res[i].status, o.status)
        assert outs[i] == o.data, i
        if o.status == 0:
            assert res[i].adler_computed == o.adler_computed
print("multi-gpu ok", L.pz_device_count())
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "multi-gpu ok" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])


# ---- the framing extension: gzip members (RFC 1952) and raw deflate --------------------------------------------------
def _framing_flags(pz, kind):
    return pz.GZIP if kind == "gzip" else pz.RAW


def test_gzip_and_raw_batch_matches_oracle(pz, oracle):
    """Every case of streams.gzip_cases() through the batch entry points (one call per framing), verdicts, messages and
    CRC-32 / Adler-32 words equal to the oracle's; valid members also against system zlib."""
    from pure_zlib_b200 import _lib
    cases = streams.gzip_cases()
    for kind in ("gzip", "raw"):
        sub = [(n, z) for n, k, z in cases if k == kind]
        for fn in (pz.zlib.decompress_batch_raw, pz.zlib.inflate_batch_raw):
            res, outs = fn([z for _, z in sub], _framing_flags(pz, kind))
            for (name, z), r, out in zip(sub, res, outs):
                o = oracle.decompress(z, framing=oracle.GZIP if kind == "gzip" else oracle.RAW)
                assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (name, r.status, r.detail, o.message)
                assert out == o.data, name
                if o.status in (0, 5):
                    assert (r.adler_computed, r.adler_stored) == (o.adler_computed, o.adler_stored), name
                if o.status != 0:
                    assert _lib.strerror(r) == o.message, name
                else:
                    assert zlib.decompressobj(31 if kind == "gzip" else -15).decompress(z) == out
    # the mirror's entry points
    g = [z for n, k, z in cases if n == "gz-fname"][0]
    assert pz.decompress_gzip(g) == pz.Right(b"hello hello hello")
    assert pz.decompress_gzip(g[:-8] + b"\0" * 8).value == pz.ChecksumError("checksum mismatch: 0 != " + format(zlib.crc32(b"hello hello hello"), "x"))
    co = zlib.compressobj(9, zlib.DEFLATED, -15)
    assert pz.decompress_raw(co.compress(b"abc" * 1000) + co.flush()) == pz.Right(b"abc" * 1000)


def test_gzip_and_raw_incremental(pz, oracle):
    """decompressIncremental behind the other framings: the event sequence of the oracle over the same chunk lists."""
    rng = np.random.default_rng(17)
    n = 0
    for name, kind, z in streams.gzip_cases():
        if len(z) < 4:
            continue
        for cuts in ([int(x) for x in rng.integers(1, len(z), 3)], list(range(0, len(z), 997))[:40]):
            pieces = _pieces(z, cuts)
            o = oracle.decompress(pieces, want_events=True, framing=oracle.GZIP if kind == "gzip" else oracle.RAW)
            st = pz.zlib.decompress_incremental(_framing_flags(pz, kind))
            events, acc, rest, err = [], b"", list(pieces), None
            while True:
                if isinstance(st, pz.NeedMore):
                    events.append((0, 0))
                    if not rest:
                        events.append((3, 0))
                        break
                    st = st.feed(rest.pop(0))
                elif isinstance(st, pz.Chunk):
                    events.append((1, len(st.data))); acc += st.data; st = st.next()
                elif isinstance(st, pz.Done):
                    events.append((2, 0)); break
                else:
                    events.append((3, 0)); err = st.error; break
            assert events == o.events, (name, cuts[:4], events[:6], o.events[:6])
            assert acc == o.data[: len(acc)] and (o.status != 0 or acc == o.data), name
            if err is not None and not (o.status == 3 and o.detail == 1):
                assert str(err) == o.message, name
            n += 1
    assert n > 30


def test_gzip_huge_stream_block_parallel(pz, oracle, huge_threshold):
    """K4 behind the gzip and raw framings (header length, 8-byte / no trailer, CRC-32 + ISIZE in K3)."""
    from pure_zlib_b200 import _lib, corpus
    L = _lib.load()
    text = corpus.text(3 << 20, 91)
    for kind, wbits in (("gzip", 31), ("raw", -15)):
        co = zlib.compressobj(9, zlib.DEFLATED, wbits)
        z = co.compress(text) + co.flush()
        bad = bytearray(z); bad[len(z) // 2] ^= 0x10
        done0 = L.pz_get_counter(1)
        res, outs = pz.zlib.inflate_batch_raw([z, bytes(bad), z[: len(z) // 3]], _framing_flags(pz, kind))
        assert L.pz_get_counter(1) - done0 >= 1, kind
        for zz, r, out in zip([z, bytes(bad), z[: len(z) // 3]], res, outs):
            o = oracle.decompress(zz, framing=oracle.GZIP if kind == "gzip" else oracle.RAW)
            assert (r.status, r.detail, r.out_len) == (o.status, o.detail, o.out_len), (kind, r.status, r.detail, o.message)
            assert out == o.data
            if o.status in (0, 5):
                assert (r.adler_computed, r.adler_stored) == (o.adler_computed, o.adler_stored), kind


def test_crc32_entry_point(pz):
    from pure_zlib_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(5)
    for n in (0, 1, 3, 511, 512, 513, 16383, 16384, 16385, 100_000, 1 << 20, (1 << 20) + 7):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert L.pz_crc32(0, d, n) == zlib.crc32(d), n
        assert L.pz_crc32(zlib.crc32(b"prefix"), d, n) == zlib.crc32(b"prefix" + d), n


# ---- O(1)-memory incremental contexts --------------------------------------------------------------------------------
def test_incremental_context_memory_is_bounded(pz):
    """A context keeps the live input (from the header of the block it is in), the last 32 KiB of output plus room for
    one launch, and a checkpoint -- not the stream (the reference's decoder is O(128 KiB), OutputWindow.hs:29-54):
    96 MiB of text through one decoder in 64 KiB pieces must never hold more than a few MiB on the device, and a
    zero run of more than 4 GiB (beyond the 32-bit counters of one launch) decodes with the right length and checksum."""
    from pure_zlib_b200 import _lib, corpus
    L = _lib.load()
    text = b"".join(corpus.text(16 << 20, 500 + k) for k in range(6))
    z = zlib.compress(text, 6)
    d = pz.zlib._Decoder()
    got = []
    adler = 1
    peak_now = 0
    for at in range(0, len(z), 1 << 16):
        piece = z[at: at + (1 << 16)]
        _lib.check(L.pz_stream_feed(d._s, piece, len(piece)), "feed")
        st = d.state()
        while isinstance(st, pz.Chunk):
            adler = zlib.adler32(st.data, adler)
            got.append(len(st.data))
            st = st.next()
        peak_now = max(peak_now, L.pz_stream_counter(d._s, _lib.PZ_SC_DEVICE_BYTES))
    assert isinstance(st, pz.Done), st
    assert sum(got) == len(text) and adler == zlib.adler32(text)
    assert all(g == 32768 for g in got[:-1])
    peak = L.pz_stream_counter(d._s, _lib.PZ_SC_DEVICE_PEAK)
    assert peak < 4 << 20, peak            # against 96 MiB of output + 30 MiB of input
    assert L.pz_stream_counter(d._s, _lib.PZ_SC_HOST_BYTES) < 8 << 20
    # more than 4 GiB from one stream: 4 GiB + 64 MiB of zeros
    total = (65 << 26)
    co = zlib.compressobj(9)
    zero = bytes(1 << 24)
    parts = [co.compress(zero) for _ in range(total >> 24)] + [co.flush()]
    zz = b"".join(parts)
    d = pz.zlib._Decoder()
    n_out, adler, ok_zero = 0, 1, True
    zeros32k = bytes(32768)
    a32k = None
    for at in range(0, len(zz), 1 << 18):
        piece = zz[at: at + (1 << 18)]
        _lib.check(L.pz_stream_feed(d._s, piece, len(piece)), "feed")
        st = d.state()
        while isinstance(st, pz.Chunk):
            n_out += len(st.data)
            ok_zero = ok_zero and (st.data == zeros32k if len(st.data) == 32768 else st.data.count(0) == len(st.data))
            st = st.next()
    assert isinstance(st, pz.Done), getattr(st, "error", st)
    assert n_out == total and ok_zero
    assert L.pz_stream_counter(d._s, _lib.PZ_SC_DEVICE_PEAK) < 160 << 20
