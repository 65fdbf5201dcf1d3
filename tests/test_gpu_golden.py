"""GPU parity against the committed golden fixtures (outputs of the compiled reference)."""
import hashlib
import json
import os

import pytest

from otezip_b200 import native
from otezip_b200.native import parse_central, default_opts

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
EXPECTED = json.load(open(os.path.join(G, "expected.json")))


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_gpu_matches_reference_golden(ctx, name):
    img = open(os.path.join(G, name + ".zip"), "rb").read()
    tab = parse_central(img)
    exp = EXPECTED[name]
    assert len(tab) == len(exp)
    if not len(tab):
        return
    out, crc, st = ctx.extract_host(img, tab, default_opts())
    L = native.Lib.get().L
    for i, e in enumerate(exp):
        acc = bool(L.otz_status_accepts(int(st[i]), 1, 1))     # otezip_verify_crc=1, reference-compatible F1
        assert acc == (e is not None), (name, i, hex(int(st[i])))
        if acc:
            data = bytes(out[int(tab["out_ofs"][i]):int(tab["out_ofs"][i]) + int(tab["uncomp_size"][i])])
            assert [len(data), int(crc[i]), hashlib.sha256(data).hexdigest()] == e, (name, i)
