"""Raw pinned-memory copy bandwidth of every rank at once (run under torchrun like bench.py): what the host side of the
box gives N GPUs that copy at the same time — the ceiling of bench.py's end-to-end number at N ranks.
  python -m torch.distributed.run --nproc-per-node N tools/copy_probe.py [GiB per copy]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from otezip_b200 import Ctx


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    dist = bench.Dist(0)
    ctx = Ctx(dist.local)
    n = int(gib * (1 << 30))
    h = ctx.pinned(n)
    h[:] = 1
    d = ctx.dev_alloc(n)
    res = {}
    for name, fn in (("h2d", lambda: ctx.h2d(d, h)), ("d2h", lambda: ctx.d2h(h, d))):
        fn()
        ctx.sync()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            fn()
        ctx.sync()
        dt = time.perf_counter() - t0
        res[name] = 4 * n / dt / 1e9
        dist.barrier()
    print("rank %d of %d: H2D %.1f GB/s, D2H %.1f GB/s (every rank copying at once)" % (dist.rank, dist.world, res["h2d"], res["d2h"]), flush=True)
    dist.barrier()


main()
