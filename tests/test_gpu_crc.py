"""CRC-32 kernel (table path, fold path, chunk combine) against zlib.crc32 == reference crc32.inc.c:40-47."""
import random
import zlib

import numpy as np
import pytest

from otezip_b200 import synth
from otezip_b200.native import parse_central, default_opts

pytestmark = pytest.mark.gpu


def test_crc_sizes_and_alignments(ctx):
    rnd = random.Random(3)
    sizes = [0, 1, 2, 15, 16, 17, 511, 512, 513, 4095, 8191, 8192, 8193, 6655, 6656, 6657, 13 * 512 * 40 - 1, 13 * 512 * 40,
             13 * 512 * 40 + 1, 600000, 1 << 20, (1 << 20) + 7]
    sizes += [rnd.randint(1, 700000) for _ in range(40)]
    ms = []
    for i, n in enumerate(sizes):
        d = synth.random_bytes(n, 100 + i)
        # names of different lengths shift the payload alignment inside the image
        ms.append(synth.Member("x" * (1 + i % 17), 0, d, n, zlib.crc32(d) & 0xFFFFFFFF, raw=d))
    img = synth.build_zip(ms)
    tab = parse_central(img)
    for verify_only in (1, 0):
        out, crc, st = ctx.extract_host(img, tab, default_opts(verify_only=verify_only))
        assert not np.count_nonzero(st), [hex(int(s)) for s in st if s]
        assert [int(c) for c in crc] == [m.crc32 for m in ms]
    # a wrong expectation must be flagged, not fatal (otezip.c:669-678)
    tab2 = tab.copy()
    tab2["crc32"][5] ^= 0x80
    out, crc, st = ctx.extract_host(img, tab2, default_opts(verify_only=1))
    assert int(st[5]) == 0x100 and int(crc[5]) == ms[5].crc32
