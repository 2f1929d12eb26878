"""ctypes binding of libpzcuda.so (include/pzcuda.h).  Loading fails loudly: there is no
fallback implementation anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PZ_LIBPZCUDA") or os.path.join(_HERE, "libpzcuda.so")  # the override is for A/B builds of the kernels

PZ_OK, PZ_ERR_HUFFMAN_TREE, PZ_ERR_FORMAT, PZ_ERR_DECOMPRESSION, PZ_ERR_HEADER, PZ_ERR_CHECKSUM, PZ_REF_BOTTOM, \
    PZ_OUTPUT_FULL, PZ_NEED_MORE = range(9)
PZ_S_NEED_MORE, PZ_S_CHUNK, PZ_S_DONE, PZ_S_ERROR = range(4)
PZ_SC_PUMPS, PZ_SC_RESUMED, PZ_SC_CKPT_BIT, PZ_SC_CKPT_BYTES, PZ_SC_DEVICE_BYTES, PZ_SC_DEVICE_PEAK, PZ_SC_HOST_BYTES = range(7)  # pz_stream_counter
PZ_F_NO_ADLER, PZ_F_COUNT_ONLY = 1, 2
PZ_F_GZIP, PZ_F_RAW = 0x20, 0x40  # framing (extension): gzip member (RFC 1952) / raw deflate (RFC 1951); default zlib
PZ_E_OK, PZ_E_CUDA, PZ_E_ARG, PZ_E_NOMEM, PZ_E_STATE = 0, -1, -2, -3, -4


PZ_MAX_DEVICES = 16


class PzConfig(C.Structure):
    """pz_config (include/pzcuda.h)."""
    _fields_ = [("device", C.c_int32), ("n_devices", C.c_int32), ("devices", C.c_int32 * PZ_MAX_DEVICES), ("reserved", C.c_int32 * 6)]


class PzResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("detail", C.c_int32), ("out_len", C.c_uint64),
                ("adler_computed", C.c_uint32), ("adler_stored", C.c_uint32), ("err_bitpos", C.c_uint64),
                ("payload", C.c_int64 * 2)]


# every symbol include/pzcuda.h declares: (name, restype, argtypes)
PZ_OPT_HUGE_BYTES = 1
PZ_F_NO_HUGE = 0x10

SYMBOLS = [
    ("pz_init", C.c_int, [C.POINTER(PzConfig)]),
    ("pz_shutdown", None, []),
    ("pz_abi_version", C.c_int, []),
    ("pz_set_option", C.c_int, [C.c_int, C.c_uint64]),
    ("pz_get_counter", C.c_uint64, [C.c_int]),
    ("pz_last_error", C.c_char_p, []),
    ("pz_inflate_batch", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(PzResult), C.c_uint32]),
    ("pz_inflate_batch_contig", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64),
                                          C.c_size_t, C.POINTER(PzResult), C.c_void_p, C.c_uint32]),
    ("pz_inflate_sizes", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(PzResult)]),
    ("pz_inflate_sizes_framed", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(PzResult), C.c_uint32]),
    ("pz_decompress_batch", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(PzResult),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32]),
    ("pz_outputs_free", None, [C.c_void_p]),
    ("pz_device_count", C.c_int, []),
    ("pz_batch_create", C.c_void_p, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t, C.c_uint32]),
    ("pz_batch_run", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("pz_batch_results", C.c_int, [C.c_void_p, C.POINTER(PzResult), C.c_void_p]),
    ("pz_batch_launches", C.c_int, [C.c_void_p]),
    ("pz_batch_destroy", None, [C.c_void_p]),
    ("pz_pinned_alloc", C.c_void_p, [C.c_size_t]),
    ("pz_pinned_free", None, [C.c_void_p]),
    ("pz_stream_new", C.c_void_p, []),
    ("pz_stream_new_framed", C.c_void_p, [C.c_uint32]),
    ("pz_stream_feed", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    ("pz_stream_next", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(PzResult)]),
    ("pz_stream_free", None, [C.c_void_p]),
    ("pz_stream_pump", C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    ("pz_stream_feed_many", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t]),
    ("pz_stream_counter", C.c_uint64, [C.c_void_p, C.c_int]),
    ("pz_strerror", C.c_size_t, [C.POINTER(PzResult), C.c_char_p, C.c_size_t]),
    ("pz_compute_code_values", C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32)]),
    ("pz_adler32", C.c_uint32, [C.c_uint32, C.c_char_p, C.c_size_t]),
    ("pz_crc32", C.c_uint32, [C.c_uint32, C.c_char_p, C.c_size_t]),
]

_lib = None


class PzCudaError(RuntimeError):
    """The library call itself failed (no device, CUDA error, bad arguments)."""


def load() -> C.CDLL:
    """dlopen libpzcuda.so and bind every declared symbol (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise PzCudaError(f"{SO_PATH} is missing: build it with `make -C pure_zlib_b200/csrc` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(L, name)          # AttributeError if the export is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc < 0:
        msg = load().pz_last_error()
        raise PzCudaError(f"{what} failed with {rc}: {msg.decode() if msg else ''}")
    return rc


def strerror(res: PzResult) -> str:
    buf = C.create_string_buffer(512)
    load().pz_strerror(C.byref(res), buf, 512)
    return buf.value.decode()
