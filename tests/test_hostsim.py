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
