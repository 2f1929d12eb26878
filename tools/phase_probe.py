#!/usr/bin/env python
"""Debug: where the service groups of K1 spend their time on a batch (needs a library built with -DPZ_PHASES:
make -C pure_zlib_b200/csrc OUT=../libpzcuda_phases.so EXTRA=-DPZ_PHASES; PZ_LIBPZCUDA=.../libpzcuda_phases.so).
Prints, per pass (sizing, decode), the share of the groups' clocks by the mode a step started in."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pure_zlib_b200 import _lib, corpus  # noqa: E402

MODES = ["IDLE (claim, begin)", "HDR (block header, tables)", "SYMS (careful symbol, block end, trailer)", "FAST", "DEAD", "WAIT (hot lane has it)", "DRAIN", "FINDRAIN"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "records4k"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
    L = _lib.load()
    raw = C.CDLL(_lib.SO_PATH)
    c = getattr(corpus, name)(n, workers=min(32, os.cpu_count() or 1))
    p64 = C.POINTER(C.c_uint64)
    d_in = torch.from_numpy(c.in_blob).cuda()
    d_out = torch.zeros(int(c.out_off[-1]) + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = (C.c_ulonglong * 16)()
    for label, flags in (("sizing", _lib.PZ_F_COUNT_ONLY), ("decode", _lib.PZ_F_NO_ADLER)):
        b = L.pz_batch_create(c.in_off.ctypes.data_as(p64), None if flags == _lib.PZ_F_COUNT_ONLY else c.out_off.ctypes.data_as(p64), c.n, flags)
        for _ in range(2):
            _lib.check(L.pz_batch_run(b, d_in.data_ptr(), None if flags == _lib.PZ_F_COUNT_ONLY else d_out.data_ptr(), st), "run")
        raw.pz_debug_phases(out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.pz_batch_run(b, d_in.data_ptr(), None if flags == _lib.PZ_F_COUNT_ONLY else d_out.data_ptr(), st), "run")
        e1.record()
        torch.cuda.synchronize()
        raw.pz_debug_phases(out)
        v = np.array(list(out), dtype=np.float64)
        groups = max(v[9], 1)
        print(f"{label}: {e0.elapsed_time(e1):.2f} ms, {int(groups)} groups, {v[8] / groups / 1e6:.2f} Mclk per group")
        for k in range(8):
            if v[k]:
                print(f"   {MODES[k]:44s} {v[k] / v[8]:6.1%}  ({v[k] / groups / 1e3:9.1f} kclk per group)")
        L.pz_batch_destroy(b)


if __name__ == "__main__":
    main()
