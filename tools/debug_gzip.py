import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import streams
import pure_zlib_b200 as pz
from pure_zlib_b200 import _lib
from pure_zlib_b200.zlib import _ptr_arrays, PzResult
L = _lib.load()
cases = [(n, z) for n, k, z in streams.gzip_cases() if k == "gzip"][:4]
zs = [z for _, z in cases]
n = len(zs)
keep, ptrs, lens = _ptr_arrays(zs)
sizes = (PzResult * n)()
print("sizes rc", L.pz_inflate_sizes_framed(ptrs, lens, n, sizes, 0x20))
for i in range(n): print(" size", cases[i][0], len(zs[i]), sizes[i].status, sizes[i].detail, sizes[i].out_len, hex(sizes[i].adler_stored), sizes[i].payload[0], sizes[i].err_bitpos)
res, outs = pz.zlib.decompress_batch_raw(zs, 0x20)
for i in range(n): print(" one-call", cases[i][0], res[i].status, res[i].detail, res[i].out_len, hex(res[i].adler_computed), hex(res[i].adler_stored))
res, outs = pz.zlib.inflate_batch_raw(zs, 0x20)
for i in range(n): print(" two-call", cases[i][0], res[i].status, res[i].detail, res[i].out_len, hex(res[i].adler_computed), hex(res[i].adler_stored))
for k in range(n):
    res, outs = pz.zlib.decompress_batch_raw([zs[k]], 0x20)
    print(" alone", cases[k][0], res[0].status, res[0].detail, res[0].out_len)
