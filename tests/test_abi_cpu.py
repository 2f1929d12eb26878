"""CPU suite, part 3: the C-ABI library loads and exports every symbol include/pzcuda.h
declares; host-side logic that needs no device.  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from pure_zlib_b200 import _lib


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pzcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pz_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    names = declared_functions()
    bound = sorted(n for n, _, _ in _lib.SYMBOLS)
    assert names == bound


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.SO_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = _lib.load()
    for name in declared_functions():
        assert hasattr(L, name), name
    assert L.pz_abi_version() == 2


def test_result_layout_matches_oracle():
    from oracle import oracle
    assert C.sizeof(_lib.PzResult) == 48 == C.sizeof(oracle.PzoResult)
    for (a, _), (b, _) in zip(_lib.PzResult._fields_, oracle.PzoResult._fields_):
        assert a == b
        assert getattr(_lib.PzResult, a).offset == getattr(oracle.PzoResult, b).offset


def test_status_codes_match_header():
    text = open(os.path.join(ROOT, "include", "pzcuda.h")).read()
    for name in ["PZ_OK", "PZ_ERR_HUFFMAN_TREE", "PZ_ERR_FORMAT", "PZ_ERR_DECOMPRESSION", "PZ_ERR_HEADER",
                 "PZ_ERR_CHECKSUM", "PZ_REF_BOTTOM", "PZ_OUTPUT_FULL", "PZ_NEED_MORE"]:
        m = re.search(name + r"\s*=\s*(\d+)", text)
        assert m and int(m.group(1)) == getattr(_lib, name)


def test_strerror_formats_reference_messages():
    L = _lib.load()
    r = _lib.PzResult()
    r.status, r.detail = 5, 1
    r.adler_stored, r.adler_computed = 0x680308B0, 0x680308B1
    assert _lib.strerror(r) == "Checksum error: checksum mismatch: 680308b0 != 680308b1"
    r.status, r.detail, r.payload[0] = 1, 3, 257
    assert _lib.strerror(r) == "Huffman tree manipulation error: Tried to add where the leaf is a node: 257"
    r.status, r.detail, r.payload[0] = 4, 2, 9
    assert _lib.strerror(r) == "Header error: Bad compression method: 9"
    assert L is not None


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pure_zlib_b200 as pz
    with pytest.raises(_lib.PzCudaError):
        pz.decompress(b"\x78\x9c\x03\x00\x00\x00\x00\x01")


def test_product_does_not_touch_the_oracle():
    """The product tree must not reference oracle/ or the host simulation."""
    pkg = os.path.join(ROOT, "pure_zlib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pz_oracle" not in src and "libpzoracle" not in src and "from oracle" not in src, f
                assert "import oracle" not in src, f


def test_config_layout_matches_header():
    """pz_config: device, n_devices, devices[PZ_MAX_DEVICES], reserved[6] (int32 each)."""
    text = open(os.path.join(ROOT, "include", "pzcuda.h")).read()
    m = re.search(r"#define PZ_MAX_DEVICES (\d+)", text)
    assert m and int(m.group(1)) == _lib.PZ_MAX_DEVICES
    assert C.sizeof(_lib.PzConfig) == 4 * (2 + _lib.PZ_MAX_DEVICES + 6)
    assert _lib.PzConfig.devices.offset == 8
