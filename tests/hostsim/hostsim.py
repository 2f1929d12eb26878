"""ctypes binding of the host-compiled decoder logic -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpzhostsim.so")


class PzResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("detail", C.c_int32), ("out_len", C.c_uint64),
                ("adler_computed", C.c_uint32), ("adler_stored", C.c_uint32), ("err_bitpos", C.c_uint64),
                ("payload", C.c_int64 * 2)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpzhostsim.so"])
        _lib = C.CDLL(_SO)
        _lib.hs_inflate.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(PzResult), C.c_int]
        _lib.hs_inflate_framed.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(PzResult), C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.hs_inflate_resume.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(PzResult), C.c_int,
                                           C.c_void_p, C.c_void_p]
        _lib.hs_fixed.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(PzResult), C.c_int, C.c_int, C.c_int]
    return _lib


def fixed(data: bytes, out_cap: int, count_only: bool = False, out_mis: int = 0, dyn: bool = False):
    """K5's (dyn: K6's) per-stream logic (pz_fixed.cuh): (completed, PzResult, bytes).  out_mis = 0: word-wide stores, 1..3: byte-wide."""
    out = C.create_string_buffer(max(out_cap, 1))
    res = PzResult()
    res.status = -1
    ok = lib().hs_fixed(bytes(data), len(data), out, out_cap, C.byref(res), int(count_only), out_mis, int(dyn))
    assert ok >= 0, "K5 wrote outside its output slice"
    return bool(ok), res, out.raw[: min(res.out_len, out_cap)] if ok else b""


def inflate(data: bytes, out_cap: int, count_only: bool = False, framing: int = 0):
    """framing: PZ_FRAME_* of pz_device.cuh (0 zlib, 1 gzip, 2 raw deflate)."""
    out = C.create_string_buffer(max(out_cap, 1))
    res = PzResult()
    lib().hs_inflate_framed(bytes(data), len(data), out, out_cap, C.byref(res), int(count_only), None, None, framing)
    return res, out.raw[: min(res.out_len, out_cap)]


class Resumable:
    """One stream decoded prefix by prefix through PzJob::resume / PzJob::ckpt: the output buffer is
    the history, the four checkpoint words of a run are the resume words of the next."""

    def __init__(self, out_cap: int):
        self.out = C.create_string_buffer(max(out_cap, 1))
        self.cap = out_cap
        self.ck = (C.c_uint32 * 4)(0, 0, 0, 0)

    def run(self, prefix: bytes):
        res = PzResult()
        rs = (C.c_uint32 * 4)(*self.ck)
        lib().hs_inflate_resume(bytes(prefix), len(prefix), self.out, self.cap, C.byref(res), 0, rs, self.ck)
        return res, self.out.raw[: min(res.out_len, self.cap)], tuple(rs)
