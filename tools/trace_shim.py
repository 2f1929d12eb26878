#!/usr/bin/env python
"""Debug: wall-clock breakdown of pz_decompress_batch on config 2 (PZ_TRACE=1 prints the library's timeline of each inner call)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pure_zlib_b200 import _lib, corpus
L = _lib.load()
c = corpus.text256k(4096, workers=16)
blobs = [bytes(c.in_blob[int(c.in_off[i]): int(c.in_off[i]) + int(c.in_len[i])]) for i in range(c.n)]
n = c.n
bufs = [C.create_string_buffer(z, len(z)) for z in blobs]
ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs]); lens = (C.c_size_t * n)(*[len(z) for z in blobs])
res = (_lib.PzResult * n)(); optrs = (C.c_void_p * n)(); h = C.c_void_p()
for k in range(4):
    t = time.perf_counter()
    _lib.check(L.pz_decompress_batch(ptrs, lens, n, res, optrs, C.byref(h), 0), "x")
    dt = time.perf_counter() - t
    L.pz_outputs_free(h)
    print(f"call {k}: {dt * 1e3:.1f} ms = {c.out_bytes / dt / 1e9:.1f} GB/s", file=sys.stderr)
