"""Debug helper: one archive with huge + odd streams through the pipelined host call (tiny sub-batches)."""
import os, sys, random, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from otezip_b200 import Ctx, synth
from otezip_b200.native import parse_central, default_opts
from tests import cases
from tests.test_gpu_inflate_twophase import _shapes
which = sys.argv[1] if len(sys.argv) > 1 else "all"
rnd = random.Random(4)
parts = {
    "shapes": lambda: _shapes(),
    "mixed": lambda: cases.mixed_archive(seed=35, n_tiny=150, n_mid=40, n_z=4, n_s=4),
    "huge": lambda: [synth.member("h0", synth.jsonlog_text(5 << 20, 1), 8), synth.member("h1", synth.jsonlog_text(3 << 20, 2), 8, level=1),
       synth.member("h2", synth.jsonlog_text(2 << 20, 4) + synth.random_bytes(600000, 5) + synth.jsonlog_text(1 << 20, 6), 8),
       synth.member("h3", b"".join(bytes([rnd.randrange(256)]) * rnd.randint(1, 2000) for _ in range(3000)), 8, strategy=zlib.Z_RLE),
       synth.member("h4", synth.jsonlog_text(2200000, 10), 8, strategy=zlib.Z_HUFFMAN_ONLY),
       synth.member("h5", synth.random_bytes(32500, 21) * 70, 8, level=9)],
}
ms = []
for k, f in parts.items():
    if which in ("all", k):
        ms += f()
img = synth.build_zip(ms)
tab = parse_central(img)
c = Ctx(0)
out, crc, st = c.extract_host(img, tab, default_opts())
print(which, "ok", len(tab), int(np.count_nonzero(st & 0xFF)))
