"""Real Zstandard (RFC 8878) frames in method-93 entries: k_zstd_tok.cuh (k_zstd_lit + k_zstd_seq + k_inflate_lz) against libzstd 1.5.5.
Parity for this kernel is NOT pinned by the reference (which rejects such frames): the frames are
made by libzstd and the decoded bytes must equal the source; the status carries the "reference rejects" flag so
that the default reference-compatible policy still agrees with the reference."""
import zlib

import numpy as np
import pytest

from otezip_b200 import native, synth
from otezip_b200.native import parse_central, default_opts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zs():
    try:
        from otezip_b200.zstdlib import Zstd
        return Zstd()
    except OSError:
        pytest.skip("libzstd.so.1 not present")


@pytest.fixture
def zctx(ctx):
    return ctx


def sources():
    out = [b"", b"a", b"hello zstd\n", b"A" * 100000, bytes(range(256)) * 40, synth.random_bytes(5000, 1), synth.random_bytes(300000, 2)]
    out += [synth.jsonlog_text(n, 50 + i) for i, n in enumerate([100, 4096, 70000, 131072, 131073, 262144, 700000])]
    out.append(synth.jsonlog_text(150000, 9) + synth.random_bytes(140000, 3) + b"\0" * 200000 + synth.jsonlog_text(90000, 10))
    return out


def test_zstd_frames_decode_bit_exact(zctx, zs, reflib):
    ctx = zctx
    ms, want = [], []
    for i, d in enumerate(sources()):
        for lvl in (1, 3, 9, 19):
            f = zs.compress(d, lvl)
            assert zs.decompress(f, len(d)) == d
            ms.append(synth.Member("z%d_%d" % (i, lvl), 93, f, len(d), zlib.crc32(d) & 0xFFFFFFFF, raw=d))
            want.append(d)
    # reference-container entries mixed in: they keep going through k_zstdref
    for i in range(4):
        d = synth.jsonlog_text(100000 + i, 200 + i)
        ms.append(synth.member("c%d" % i, d, 93))
        want.append(d)
    img = synth.build_zip(ms)
    tab = parse_central(img)
    out, crc, st = ctx.extract_host(img, tab, default_opts())
    L = native.Lib.get().L
    for i, (m, d) in enumerate(zip(ms, want)):
        s = int(st[i])
        assert (s & 0xFF) == 0, (m.name, hex(s))
        assert not (s & native.STF_CRC_MISMATCH), m.name
        real = m.name.startswith("z")
        assert bool(s & native.STF_REF_EOB) == real, (m.name, hex(s))      # real frames: "the reference rejects this"
        assert bool(L.otz_status_accepts(s, 1, 0)) and bool(L.otz_status_accepts(s, 1, 1)) == (not real)
        o = int(tab["out_ofs"][i])
        assert bytes(out[o:o + len(d)]) == d, m.name
        assert int(crc[i]) == m.crc32
    # and the compiled reference indeed rejects every real frame while reading its own container
    err, got = reflib.extract_bytes(img, verify_crc=1)
    assert err == 0
    for m, g, d in zip(ms, got, want):
        assert (g is None) == m.name.startswith("z") and (g is None or g == d)


def test_zstd_frame_structure_variants(zctx, zs):
    """What the walk shared by k_zstd_lit and k_zstd_seq (zs_walk) has to get through: several frames per entry, skippable
    frames between them, frames with a content checksum, frames without Frame_Content_Size (window descriptor instead of
    the single-segment flag), small windows, and entries made of raw / RLE blocks only."""
    import struct
    a, b, c = synth.jsonlog_text(200000, 21), synth.jsonlog_text(70000, 22), synth.random_bytes(3000, 23)
    skip = struct.pack("<II", 0x184D2A53, 11) + b"hello world"
    cases = {
        "two_frames": (zs.compress(a, 3) + zs.compress(b, 9), a + b),
        "skippable": (skip + zs.compress(a, 1) + skip + zs.compress(c, 3) + skip, a + c),
        "checksum": (zs.compress_adv(a, 3, checksum=True), a),
        "checksum_x2": (zs.compress_adv(a, 5, checksum=True) + zs.compress_adv(b, 1, checksum=True), a + b),
        "no_fcs": (zs.compress_adv(a, 3, content_size=False), a),
        "no_fcs_cksum_w17": (zs.compress_adv(a + b, 7, checksum=True, content_size=False, window_log=17), a + b),
        "raw_blocks": (zs.compress(synth.random_bytes(300000, 24), 3), synth.random_bytes(300000, 24)),
        "rle_blocks": (zs.compress(b"\x07" * 500000, 3), b"\x07" * 500000),
        "rle_then_text": (zs.compress(b"z" * 140000 + a, 3), b"z" * 140000 + a),
        "empty_then_text": (zs.compress(b"", 3) + zs.compress(b, 3), b),
    }
    ms, want = [], []
    for name, (f, d) in cases.items():
        assert zs.decompress(f, len(d)) == d or name in ("two_frames", "skippable", "checksum_x2", "empty_then_text")   # (ZSTD_decompress reads all frames; kept loose)
        ms.append(synth.Member(name, 93, f, len(d), zlib.crc32(d) & 0xFFFFFFFF, raw=d))
        want.append(d)
    img = synth.build_zip(ms)
    tab = parse_central(img)
    out, crc, st = zctx.extract_host(img, tab, default_opts())
    for i, (m, d) in enumerate(zip(ms, want)):
        s = int(st[i])
        assert (s & 0xFF) == 0 and not (s & native.STF_CRC_MISMATCH), (m.name, hex(s))
        o = int(tab["out_ofs"][i])
        assert bytes(out[o:o + len(d)]) == d, m.name
        assert int(crc[i]) == m.crc32
    # a frame cut inside its checksum, and a declared Frame_Content_Size that is wrong: never a clean success
    bad = [zs.compress_adv(a, 3, checksum=True)[:-2], bytearray(zs.compress(b, 3))]
    bad[1][5] ^= 0x10   # (byte 5 belongs to Frame_Content_Size in a single-segment frame of this size)
    ms = [synth.Member("bad%d" % i, 93, bytes(f), len(d), zlib.crc32(d) & 0xFFFFFFFF) for i, (f, d) in enumerate(zip(bad, (a, b)))]
    img = synth.build_zip(ms)
    tab = parse_central(img)
    out, crc, st = zctx.extract_host(img, tab, default_opts())
    for i in range(2):
        assert (int(st[i]) & 0xFF) != 0, (i, hex(int(st[i])))


def test_corrupt_zstd_frames_are_rejected(zctx, zs):
    ctx = zctx
    d = synth.jsonlog_text(200000, 77)
    f = zs.compress(d, 3)
    ms = []
    import random
    rnd = random.Random(4)
    for k in range(64):
        b = bytearray(f)
        pos = rnd.randrange(4, len(b))
        b[pos] ^= 1 << rnd.randrange(8)
        ms.append(synth.Member("x%d" % k, 93, bytes(b), len(d), zlib.crc32(d) & 0xFFFFFFFF))
    ms.append(synth.Member("trunc", 93, f[:len(f) // 2], len(d), zlib.crc32(d) & 0xFFFFFFFF))
    img = synth.build_zip(ms)
    tab = parse_central(img)
    out, crc, st = ctx.extract_host(img, tab, default_opts())      # must terminate and never report a clean success
    for i in range(len(ms)):
        s = int(st[i])
        ok = (s & 0xFF) == 0 and not (s & native.STF_CRC_MISMATCH)
        if ok:   # a flip in an unused header bit may leave the data intact
            o = int(tab["out_ofs"][i])
            assert bytes(out[o:o + len(d)]) == d


def test_zstd_through_libzip_api_policy(zs, tmp_path):
    """zip_fopen_index: the reference rejects real Zstandard frames, so the default (otezip_ref_compat = 1) returns
    NULL for them exactly like the reference; otezip_ref_compat = 0 hands out the decoded bytes."""
    from otezip_b200.zipapi import ZipApi
    api = ZipApi()
    d = synth.jsonlog_text(300000, 5)
    ms = [synth.Member("real.zst", 93, zs.compress(d, 3), len(d), zlib.crc32(d) & 0xFFFFFFFF), synth.member("container", d, 93)]
    p = tmp_path / "z.zip"
    p.write_bytes(synth.build_zip(ms))
    assert api.read_all(str(p))[2] == [None, d]
    api.ref_compat.value = 0
    try:
        assert api.read_all(str(p))[2] == [d, d]
    finally:
        api.ref_compat.value = 1
