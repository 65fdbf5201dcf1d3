"""GPU parity: the CUDA batch extract (through the C-ABI) vs the CPU oracle on the same inputs."""
import numpy as np
import pytest

from otezip_b200 import native, synth
from otezip_b200.native import parse_central, default_opts
from tests import cases

pytestmark = pytest.mark.gpu


def check_against_oracle(ctx, oracle, img, members=None, verify_only=0):
    tab = parse_central(img)
    out, crc, st = ctx.extract_host(img, tab, default_opts(verify_only=verify_only))
    rc, oents = oracle.load_central(img)
    assert rc == 0 and len(oents) == len(tab)
    ost, ocrc, oout, oofs = oracle.extract_all(img, oents)     # verify_crc=1
    L = native.Lib.get().L
    n_acc = 0
    for i in range(len(tab)):
        acc = L.otz_status_accepts(int(st[i]), 1, 1)
        assert bool(acc) == (ost[i] == 0), (i, hex(int(st[i])), int(ost[i]))
        if acc:
            n_acc += 1
            assert int(crc[i]) == int(ocrc[i]), i
            if not (verify_only and tab["method"][i] == 0):
                a = out[int(tab["out_ofs"][i]):int(tab["out_ofs"][i]) + int(tab["uncomp_size"][i])]
                b = oout[int(oofs[i]):int(oofs[i]) + int(tab["uncomp_size"][i])]
                assert np.array_equal(a, b), i
    return n_acc, st


def test_mixed_archive_matches_oracle(ctx, oracle):
    ms = cases.mixed_archive()
    img = synth.build_zip(ms)
    n_acc, st = check_against_oracle(ctx, oracle, img)
    assert n_acc > 400
    # the F1 flag must fire for some tiny entries and those are valid RFC 1951 streams
    flagged = [i for i in range(len(ms)) if (int(st[i]) & 0xFF) == 0 and int(st[i]) & native.STF_REF_EOB]
    assert len(flagged) > 0


def test_store_verify_only(ctx, oracle):
    ms = synth.config_c2(24, size=1 << 20) + [synth.member("e", b"", 0), synth.member("x", b"abc", 0)]
    img = synth.build_zip(ms)
    n_acc, _ = check_against_oracle(ctx, oracle, img, verify_only=1)
    assert n_acc == len(ms)
