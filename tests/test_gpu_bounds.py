"""Software bounds checks (the pool's GPUs do not admit compute-sanitizer): the hardest shapes of the parity suite run
against libotezip_b200_dbg.so — the same sources with -DOTZ_BOUNDS_CHECK, every unmasked / per-lane indexed access of the
speculative tokenizer, the LZ executor, the Zstandard tokenizers and the compressor's second search pass guarded by a
counter — in a child process; all counters must stay zero and
the results must still be the oracle's."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "otezip_b200", "libotezip_b200_dbg.so")

CHILD = r'''
import ctypes as C, json, os, sys, random, zlib
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import Oracle
from otezip_b200 import Ctx, synth
from otezip_b200.native import Lib, parse_central, default_opts
from tests import cases
from tests.test_gpu_inflate_twophase import _shapes
L = Lib.get().L
L.otz_debug_violations.argtypes = [C.c_void_p, C.c_int]
rnd = random.Random(4)
ms = _shapes() + cases.mixed_archive(seed=35, n_tiny=150, n_mid=40, n_z=4, n_s=4)
ms += [synth.member("h0", synth.jsonlog_text(5 << 20, 1), 8), synth.member("h1", synth.jsonlog_text(3 << 20, 2), 8, level=1),
       synth.member("h2", synth.jsonlog_text(2 << 20, 4) + synth.random_bytes(600000, 5) + synth.jsonlog_text(1 << 20, 6), 8),
       synth.member("h3", b"".join(bytes([rnd.randrange(256)]) * rnd.randint(1, 2000) for _ in range(3000)), 8, strategy=zlib.Z_RLE),
       synth.member("h4", synth.jsonlog_text(2200000, 10), 8, strategy=zlib.Z_HUFFMAN_ONLY),
       synth.member("h5", synth.random_bytes(32500, 21) * 70, 8, level=9)]
try:   # real Zstandard frames (the oracle rejects them like the reference does; here only the counters matter)
    from otezip_b200.zstdlib import Zstd
    zs = Zstd()
    for i, (d, lvl) in enumerate([(synth.jsonlog_text(300000, 31), 1), (synth.jsonlog_text(262144, 32), 3), (synth.jsonlog_text(150000, 33), 19),
                                  (synth.random_bytes(200000, 34), 3), (b"q" * 300000 + synth.jsonlog_text(5000, 35), 3), (b"", 3)]):
        f = zs.compress(d, lvl) if i != 1 else zs.compress(d[:100000], lvl) + zs.compress_adv(d[100000:], 5, checksum=True, content_size=False)
        ms.append(synth.Member("zs%%d" %% i, 93, f, len(d), zlib.crc32(d) & 0xFFFFFFFF, raw=d))
except OSError:
    pass
img = synth.build_zip(ms)
tab = parse_central(img)
o = Oracle()
rc, oents = o.load_central(img)
ost, ocrc, oout, oofs = o.extract_all(img, oents)
bad = 0
for env in ({}, {"OTZ_SEG_PAR_RING": "8192"}, {"OTZ_LZ_RING": "16384", "OTZ_SPEC_GRID": "3"}, {"OTZ_PIPE_BYTES": "400000"}):
    os.environ.update(env)
    c = Ctx(0)
    out, crc, st = c.extract_host(img, tab, default_opts())
    c.close()
    for k in env:
        del os.environ[k]
    for i in range(len(tab)):
        ok = (int(st[i]) & 0xFF) == 0 and not (int(st[i]) & 0x300)
        if ok != (ost[i] == 0):
            bad += 1
        elif ok:
            n = int(tab["uncomp_size"][i])
            if not np.array_equal(out[int(tab["out_ofs"][i]):int(tab["out_ofs"][i]) + n], oout[int(oofs[i]):int(oofs[i]) + n]):
                bad += 1
# the write path: both compression levels, every stream through zlib
c = Ctx(0)
src = [synth.jsonlog_text(n_, 40 + i) for i, n_ in enumerate([70000, 262144, 65280, 65281, 5])] + [synth.random_bytes(100000, 41), b"A" * 200000]
for meth in (8, 8 | 0x100):
    for s_, (m, p, crc_) in zip(src, c.deflate_host(src, [meth] * len(src))):
        if (zlib.decompress(p, -15) if m == 8 else p) != s_ or crc_ != (zlib.crc32(s_) & 0xFFFFFFFF):
            bad += 1
c.close()
v = (C.c_uint64 * 32)()
n = L.otz_debug_violations(v, 32)
print(json.dumps({"slots": n, "violations": [int(x) for x in v[:max(n, 0)]], "mismatches": bad, "entries": len(tab)}))
'''


@pytest.mark.skipif(not os.path.exists(DBG), reason="debug library not built (make debug)")
def test_bounds_checked_build_counts_no_violation():
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, timeout=1200, cwd=ROOT,
                       env=dict(os.environ, OTEZIP_B200_LIB=DBG))
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["slots"] > 0, "the debug library must have the checks compiled in"
    assert d["mismatches"] == 0 and sum(d["violations"]) == 0, d
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(d, open(os.path.join(out, "bounds_check.json"), "w"))


def test_release_build_has_no_checks_compiled_in():
    import ctypes as C
    from otezip_b200.native import Lib
    L = Lib.get().L
    L.otz_debug_violations.argtypes = [C.c_void_p, C.c_int]
    if os.environ.get("OTEZIP_B200_LIB"):
        pytest.skip("running against another build")
    assert L.otz_debug_violations(None, 0) == -1
