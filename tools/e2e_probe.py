"""Times otz_extract_host (C-ABI, host buffers) call by call on a bench workload and prints the fallback counter."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from otezip_b200 import Ctx


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n_entries = int(sys.argv[2]) if len(sys.argv) > 2 else None
    ctx = Ctx(0)
    wl = bench.workload(name, 0, n_entries, ctx.pinned)
    img, tab, n = wl["image"], wl["table"], len(wl["table"])
    out = ctx.pinned(wl["out_bytes"])
    crc = np.zeros(n, dtype=np.uint32)
    st = np.zeros(n, dtype=np.int32)
    L = ctx.L
    for i in range(5):
        t0 = time.perf_counter()
        ctx.lib.check(L.otz_extract_host(ctx.h, img.ctypes.data_as(C.c_void_p), img.nbytes, tab.ctypes.data_as(C.c_void_p), n,
                                         C.byref(wl["opts"]), out.ctypes.data_as(C.c_void_p), wl["out_bytes"],
                                         crc.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p)), "otz_extract_host")
        ctx.sync()
        dt = time.perf_counter() - t0
        print("call %d: %.1f ms, fallbacks %d, bad status %d" % (i, dt * 1e3, int(L.otz_inflate_fallbacks(ctx.h)), int(np.count_nonzero(st & 0xFF))))


main()
