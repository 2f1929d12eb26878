#!/usr/bin/env python
"""Debug: one stream (hex on the command line, framing 0/1/2) through both batch entry points, next to the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle
import pure_zlib_b200 as pz
from pure_zlib_b200 import _lib
fr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
flags = {0: 0, 1: _lib.PZ_F_GZIP, 2: _lib.PZ_F_RAW}[fr]
for hx in sys.argv[1].split(","):
    z = bytes.fromhex(hx)
    o = oracle.decompress(z, framing=fr)
    print("oracle ", o.status, o.detail, o.out_len, hex(o.adler_computed), hex(o.adler_stored), o.message)
    for fn in (pz.zlib.decompress_batch_raw, pz.zlib.inflate_batch_raw):
        for reps in (1, 9000):
            res, outs = fn([z] * reps, flags)
            r = res[reps // 2]
            print(fn.__name__, reps, r.status, r.detail, r.out_len, hex(r.adler_computed), hex(r.adler_stored), _lib.strerror(r), outs[reps // 2] == o.data)
