"""Every (lanes per stream, ring size) instantiation of the inflate kernel must give identical results."""
import os

import numpy as np
import pytest

from otezip_b200 import Ctx, synth
from otezip_b200.native import parse_central, default_opts
from tests import cases

pytestmark = pytest.mark.gpu
CONFIGS = [(32, 16384), (32, 4096), (32, 2048), (16, 4096), (16, 2048), (8, 4096), (8, 2048), (8, 1024), (4, 2048), (4, 1024)]


@pytest.fixture(scope="module")
def archive():
    ms = cases.mixed_archive(seed=21, n_tiny=150, n_mid=40, n_z=0, n_s=0)
    # long-distance and long-run matches: exercise the far (HBM) back-reference path and overlapping copies
    blob = synth.random_bytes(30000, 5)
    ms.append(synth.member("far", blob + blob + blob[:5000], 8, level=9))
    ms.append(synth.member("runs", b"ab" * 40000 + b"x" * 70000 + bytes(range(256)) * 100, 8, level=9))
    ms.append(synth.member("big", synth.jsonlog_text(3 << 20, 77), 8, level=6))
    return synth.build_zip(ms)


@pytest.mark.parametrize("g,w", CONFIGS)
def test_config_matches_oracle(g, w, archive, oracle):
    os.environ["OTZ_INFLATE_TILE"], os.environ["OTZ_INFLATE_RING"] = str(g), str(w)
    try:
        c = Ctx(0)
    finally:
        del os.environ["OTZ_INFLATE_TILE"], os.environ["OTZ_INFLATE_RING"]
    tab = parse_central(archive)
    out, crc, st = c.extract_host(archive, tab, default_opts())
    c.close()
    rc, oents = oracle.load_central(archive)
    ost, ocrc, oout, oofs = oracle.extract_all(archive, oents)
    for i in range(len(tab)):
        ok = (int(st[i]) & 0xFF) == 0 and not (int(st[i]) & 0x300)
        assert ok == (ost[i] == 0), (i, hex(int(st[i])), int(ost[i]))
        if ok:
            n = int(tab["uncomp_size"][i])
            assert np.array_equal(out[int(tab["out_ofs"][i]):int(tab["out_ofs"][i]) + n], oout[int(oofs[i]):int(oofs[i]) + n]), i
