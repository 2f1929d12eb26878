"""ctypes binding of the parity oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (see oracle/pz_oracle.h).  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpzoracle.so")


class PzoResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("detail", C.c_int32),
        ("out_len", C.c_uint64),
        ("adler_computed", C.c_uint32),
        ("adler_stored", C.c_uint32),
        ("err_bitpos", C.c_uint64),
        ("payload", C.c_int64 * 2),
    ]


class PzoEvent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("len", C.c_uint64)]


EV_NEED_MORE, EV_CHUNK, EV_DONE, EV_ERROR = 0, 1, 2, 3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpzoracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.pzo_decompress.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(PzoResult), C.POINTER(PzoEvent), C.c_size_t,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.pzo_decompress.restype = C.c_int
        L.pzo_decompress_framed.argtypes = L.pzo_decompress.argtypes + [C.c_int]
        L.pzo_decompress_framed.restype = C.c_int
        L.pzo_compute_code_values.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32)]
        L.pzo_compute_code_values.restype = C.c_int
        L.pzo_tree_check.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
        L.pzo_tree_check.restype = C.c_int
        L.pzo_adler32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.pzo_adler32.restype = C.c_uint32
        L.pzo_strerror.argtypes = [C.POINTER(PzoResult), C.c_char_p, C.c_size_t]
        L.pzo_strerror.restype = C.c_size_t
        _lib = L
    return _lib


@dataclass
class Verdict:
    """What `decompress` returns, in comparable form."""
    status: int
    detail: int
    data: bytes              # decoded bytes (complete for status 0; prefix decoded so far otherwise)
    out_len: int
    adler_computed: int
    adler_stored: int
    payload: tuple
    message: str             # `show` of the Left value; "" for Right
    events: list = field(default_factory=list)   # [(kind, len)]
    published: int = 0

    @property
    def ok(self) -> bool:
        return self.status == 0

    def key(self):
        """The fields that define parity (bytes + verdict)."""
        if self.status == 0:
            return (0, self.data, self.adler_computed)
        return (self.status, self.detail, self.message)


ZLIB, GZIP, RAW = 0, 1, 2


def decompress(chunks, out_cap: int | None = None, want_events: bool = False, framing: int = ZLIB) -> Verdict:
    """`Codec.Compression.Zlib.decompress` on a lazy ByteString made of `chunks`
    (bytes or a list of bytes).  framing = GZIP / RAW: the extension of pz_oracle.h (pzo_decompress_framed)."""
    if isinstance(chunks, (bytes, bytearray, memoryview)):
        chunks = [bytes(chunks)]
    chunks = [bytes(c) for c in chunks]
    blob = b"".join(chunks)
    n = len(chunks)
    lens = (C.c_size_t * max(n, 1))(*[len(c) for c in chunks])
    if out_cap is None:
        out_cap = max(1 << 16, len(blob) * 4)
    L = lib()
    while True:
        out = C.create_string_buffer(max(out_cap, 1))
        res = PzoResult()
        ev_cap = 1 << 16 if want_events else 0
        ev = (PzoEvent * max(ev_cap, 1))()
        n_ev = C.c_size_t(0)
        pub = C.c_uint64(0)
        inbuf = C.create_string_buffer(blob, max(len(blob), 1))
        L.pzo_decompress_framed(inbuf, lens, n, out, out_cap, C.byref(res), ev if want_events else None, ev_cap,
                                C.byref(n_ev), C.byref(pub), framing)
        if res.out_len <= out_cap:
            break
        out_cap = int(res.out_len)
    msg = C.create_string_buffer(512)
    L.pzo_strerror(C.byref(res), msg, 512)
    events = [(ev[i].kind, int(ev[i].len)) for i in range(min(n_ev.value, ev_cap))] if want_events else []
    return Verdict(res.status, res.detail, out.raw[: res.out_len], int(res.out_len), res.adler_computed,
                   res.adler_stored, (int(res.payload[0]), int(res.payload[1])), msg.value.decode(), events,
                   int(pub.value))


def compute_code_values(pairs):
    n = len(pairs)
    sym = (C.c_int32 * max(n, 1))(*[p[0] for p in pairs])
    ln = (C.c_int32 * max(n, 1))(*[p[1] for p in pairs])
    out = (C.c_int32 * (3 * max(n, 1)))()
    m = lib().pzo_compute_code_values(sym, ln, n, out)
    return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(m)]


def tree_check(lens: bytes):
    v = C.c_int64(0)
    e = lib().pzo_tree_check(bytes(lens), len(lens), C.byref(v))
    return e, int(v.value)


def adler32(data: bytes, init: int = 1) -> int:
    return lib().pzo_adler32(init, bytes(data), len(data))
