"""Two-phase inflate (k_inflate_spec: warp-per-stream / CTA-per-huge-stream speculative entropy decode; k_inflate_lz:
warp-per-stream LZ77 execution; k_inflate as the fallback for declined streams) against the oracle and against the
one-kernel decoder."""
import os
import random
import zlib

import numpy as np
import pytest

from otezip_b200 import Ctx, synth
from otezip_b200.native import parse_central, default_opts
from tests import cases

pytestmark = pytest.mark.gpu


def _ctx(mode=None, ring=None):
    env = {}
    if mode:
        env["OTZ_INFLATE_MODE"] = mode
    if ring:
        env["OTZ_LZ_RING"] = str(ring)
    os.environ.update(env)
    try:
        return Ctx(0)
    finally:
        for k in env:
            del os.environ[k]


def _check(img, oracle, c, expect_fallbacks=None):
    tab = parse_central(img)
    out, crc, st = c.extract_host(img, tab, default_opts())
    fb = int(c.L.otz_inflate_fallbacks(c.h))
    rc, oents = oracle.load_central(img)
    ost, ocrc, oout, oofs = oracle.extract_all(img, oents)
    for i in range(len(tab)):
        ok = (int(st[i]) & 0xFF) == 0 and not (int(st[i]) & 0x300)
        assert ok == (ost[i] == 0), (i, hex(int(st[i])), int(ost[i]))
        if ok:
            n = int(tab["uncomp_size"][i])
            a = out[int(tab["out_ofs"][i]):int(tab["out_ofs"][i]) + n]
            b = oout[int(oofs[i]):int(oofs[i]) + n]
            assert np.array_equal(a, b), (i, n, int(np.argmax(a != b)) if n else -1)
            assert int(crc[i]) == int(ocrc[i]), i
    if expect_fallbacks is not None:
        assert fb == expect_fallbacks, fb
    return fb, st, out


def _shapes():
    rnd = random.Random(5)
    ms = []
    blob = synth.random_bytes(30000, 5)
    ms.append(synth.member("far", blob + blob + blob[:5000], 8, level=9))                      # 30000-byte distances
    ms.append(synth.member("runs", b"ab" * 40000 + b"x" * 70000 + bytes(range(256)) * 100, 8, level=9))   # overlapping copies
    ms.append(synth.member("big", synth.jsonlog_text(3 << 20, 77), 8, level=6))               # many blocks
    ms.append(synth.member("lits", synth.random_bytes(70000, 9), 8, strategy=zlib.Z_HUFFMAN_ONLY))  # literal runs > 511
    ms.append(synth.member("fixed", synth.jsonlog_text(50000, 3), 8, strategy=zlib.Z_FIXED))
    ms.append(synth.member("rle", b"".join(bytes([rnd.randrange(256)]) * rnd.randint(1, 600) for _ in range(500)), 8, strategy=zlib.Z_RLE))
    ms.append(synth.member("flush", synth.jsonlog_text(300000, 4), 8, full_flush_every=4096))   # empty stored blocks mid-stream
    ms.append(synth.member("empty", b"", 8))
    ms.append(synth.member("one", b"x", 8))
    ms.append(synth.member("stored", synth.random_bytes(5000, 1), 8, level=0))                 # stored payload -> fallback
    for i in range(40):
        ms.append(synth.member("j%d" % i, synth.jsonlog_text(rnd.randint(1, 90000), 100 + i), 8, level=rnd.choice([1, 6, 9]),
                               ref_safe=bool(i & 1)))
    # binary-like data: long literal/length codes -> second-level tables
    for i in range(8):
        d = bytes((rnd.randrange(256) if rnd.random() < 0.7 else 0) for _ in range(rnd.randint(3000, 60000)))
        ms.append(synth.member("b%d" % i, d, 8))
    return ms


def test_shapes_match_oracle(oracle):
    img = synth.build_zip(_shapes())
    c = _ctx()
    fb, st, out = _check(img, oracle, c)
    c.close()
    assert fb <= 3, fb   # only the stored-payload stream (and nothing structural) may take the fallback


@pytest.mark.parametrize("ring", [4096, 8192, 16384])
def test_rings(ring, oracle):
    img = synth.build_zip(_shapes())
    c = _ctx(ring=ring)
    _check(img, oracle, c)
    c.close()


def test_mixed_archive_with_errors(oracle):
    # tiny entries (many trip the reference's end-of-block rule), every strategy/level, stored blocks
    ms = cases.mixed_archive(seed=33, n_tiny=300, n_mid=60, n_z=5, n_s=5)
    img = synth.build_zip(ms)
    c = _ctx()
    _check(img, oracle, c)
    c.close()


def test_same_status_words_as_one_kernel_decoder():
    ms = cases.mixed_archive(seed=34, n_tiny=200, n_mid=40, n_z=0, n_s=0)
    img = bytearray(synth.build_zip(ms))
    tab = parse_central(bytes(img))
    # corrupt a few payloads: both decoders must report the same status words and produce the same bytes where accepted
    rnd = random.Random(3)
    for _ in range(25):
        i = rnd.randrange(len(tab))
        if int(tab["comp_size"][i]) > 8:
            lfh = int(tab["lfh_ofs"][i])
            img[lfh + 30 + len(ms[i].name) + rnd.randrange(int(tab["comp_size"][i]))] ^= 1 << rnd.randrange(8)
    img = bytes(img)
    a = _ctx()
    out_a, crc_a, st_a = a.extract_host(img, tab, default_opts())
    a.close()
    b = _ctx(mode="legacy")
    out_b, crc_b, st_b = b.extract_host(img, tab, default_opts())
    b.close()
    assert np.array_equal(st_a, st_b)
    for i in range(len(tab)):
        if (int(st_a[i]) & 0xFF) == 0:
            n, o = int(tab["uncomp_size"][i]), int(tab["out_ofs"][i])
            assert np.array_equal(out_a[o:o + n], out_b[o:o + n]), i
            assert int(crc_a[i]) == int(crc_b[i])


def test_many_uniform_streams_no_fallback(oracle):
    pool = synth.TextPool(4 << 20, 11)
    ms = [synth.member("e%d" % i, pool.take(65536), 8) for i in range(600)]
    img = synth.build_zip(ms)
    c = _ctx()
    _check(img, oracle, c, expect_fallbacks=0)
    c.close()


@pytest.mark.parametrize("seg_exec", [None, "serial", "limit", "ring8192"])
@pytest.mark.parametrize("seg_grid", [None, "2"])
def test_huge_streams_segmented(oracle, seg_grid, seg_exec):
    """Entries >= 2 MiB: one 4-warp CTA per stream (k_inflate_spec<4>: 128 pieces per round), token stream cut into
    segments, chain check (k_seg_stitch) against the oracle, in shapes that stress it: many blocks, stored and fixed
    blocks in between, full-flush points, an incompressible middle, a stream that is one single block.
    seg_exec: how the accepted chains are executed — every segment by its own warp over 16-bit symbols with markers for
    the 32 KiB before it (k_inflate_lz<.., PAR> + k_seg_window + k_seg_translate; default), one warp walking the chain
    ("serial"), or a symbol buffer that only has room for some of the streams ("limit": both executors in one run)."""
    rnd = random.Random(9)
    ms = [synth.member("h0", synth.jsonlog_text(5 << 20, 1), 8),
          synth.member("h1", synth.jsonlog_text(3 << 20, 2), 8, level=1),
          synth.member("h2", synth.jsonlog_text(3 << 20, 3), 8, level=9, ref_safe=False),
          synth.member("h3", synth.jsonlog_text(2 << 20, 4) + synth.random_bytes(600000, 5) + synth.jsonlog_text(1 << 20, 6), 8),
          synth.member("h4", synth.jsonlog_text(4 << 20, 7), 8, full_flush_every=300000),
          synth.member("h5", synth.jsonlog_text(2500000, 8), 8, strategy=zlib.Z_FIXED),
          synth.member("h6", b"".join(bytes([rnd.randrange(256)]) * rnd.randint(1, 2000) for _ in range(3000)), 8, strategy=zlib.Z_RLE),
          synth.member("h7", synth.jsonlog_text(2200000, 10), 8, strategy=zlib.Z_HUFFMAN_ONLY)]
    ms += [synth.member("s%d" % i, synth.jsonlog_text(rnd.randint(1000, 300000), 20 + i), 8) for i in range(30)]
    img = synth.build_zip(ms)
    if seg_grid:   # two CTAs for all streams: every group decodes many streams one after the other
        os.environ["OTZ_SPEC_GRID"] = seg_grid
    if seg_exec == "serial":
        os.environ["OTZ_SEG_EXEC"] = "serial"
    elif seg_exec == "limit":
        os.environ["OTZ_SEG_SYM_LIMIT"] = str(9 << 20)
    elif seg_exec in ("ring8192", "ring4096"):
        os.environ["OTZ_SEG_PAR_RING"] = seg_exec[4:]
    try:
        c = _ctx()
        fb, st, out = _check(img, oracle, c)
        c.close()
    finally:
        for k in ("OTZ_SPEC_GRID", "OTZ_SEG_EXEC", "OTZ_SEG_SYM_LIMIT", "OTZ_SEG_PAR_RING"):
            os.environ.pop(k, None)
    assert fb <= 3, fb   # (the incompressible middle of h3 is stored blocks with payload: k_inflate takes that stream)


def test_corrupt_huge_streams_same_status_as_one_kernel_decoder():
    """Huge entries with flipped bits / cut tails / a planted fake block header: the segmented decode must end with
    exactly the status words and bytes of the one-kernel decoder (it hands every stream whose chain does not close
    to that decoder)."""
    rnd = random.Random(12)
    base = synth.jsonlog_text(3 << 20, 31)
    good = synth.deflate_raw(base, 6)
    ms = []
    for i in range(10):
        b = bytearray(good)
        for _ in range(1 + i % 3):
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        ms.append(synth.Member("flip%d" % i, 8, bytes(b), len(base), zlib.crc32(base) & 0xFFFFFFFF))
    for i in range(4):
        ms.append(synth.Member("cut%d" % i, 8, good[:len(good) - rnd.randrange(1, len(good) // 2)], len(base), zlib.crc32(base) & 0xFFFFFFFF))
    ms.append(synth.Member("wrongsize", 8, good, len(base) - 5, zlib.crc32(base) & 0xFFFFFFFF))
    ms.append(synth.Member("short", 8, good, len(base) + 100, zlib.crc32(base + b"\0" * 100) & 0xFFFFFFFF))
    # a valid dynamic-block header (copied from the stream's own first block) planted inside incompressible data:
    # the search finds it, the chain must not follow it
    noise = synth.random_bytes(200000, 6)
    planted = noise[:100000] + good[:4000] + noise[100000:]
    d2 = base[:1 << 20] + planted + base[1 << 20:]
    ms.append(synth.member("planted", d2, 8))
    ms.append(synth.member("ok", base, 8))
    img = synth.build_zip(ms)
    tab = parse_central(img)
    a = _ctx()
    out_a, crc_a, st_a = a.extract_host(img, tab, default_opts())
    a.close()
    b = _ctx(mode="legacy")
    out_b, crc_b, st_b = b.extract_host(img, tab, default_opts())
    b.close()
    assert np.array_equal(st_a, st_b), (list(map(hex, st_a)), list(map(hex, st_b)))
    n_ok = 0
    for i in range(len(tab)):
        if (int(st_a[i]) & 0xFF) == 0:
            n, o = int(tab["uncomp_size"][i]), int(tab["out_ofs"][i])
            assert np.array_equal(out_a[o:o + n], out_b[o:o + n]), i
            assert int(crc_a[i]) == int(crc_b[i])
            n_ok += 1
    assert n_ok >= 3


@pytest.mark.parametrize("huge_bytes", [None, "60000"])
def test_random_encoder_settings_match_oracle(oracle, huge_bytes):
    """Streams from every corner of zlib's parameter space — window 512 B..32 KiB, memLevel 1..9 (memLevel 1 ends a block
    every 128 symbols: thousands of block headers per entry), all levels and strategies, sync / full flushes at random
    points — through the lane-per-stream path, against the oracle.  huge_bytes = 60000 sends the larger ones through the
    segmented decode and the parallel segment execution instead (hundreds of tiny segments per stream)."""
    rnd = random.Random(77)
    ms = []
    for i in range(160):
        n = rnd.choice([0, 1, 17, 300, 5000, 40000, 150000, 400000])
        kind = rnd.randrange(4)
        if kind == 0:
            d = synth.jsonlog_text(n, 300 + i)
        elif kind == 1:
            d = synth.random_bytes(n, i)
        elif kind == 2:
            d = bytes((rnd.randrange(256) if rnd.random() < 0.3 else 65) for _ in range(n))
        else:
            d = (synth.jsonlog_text(max(n // 3, 1), i) * 3)[:n]
        c = zlib.compressobj(rnd.choice([1, 2, 4, 6, 9]), zlib.DEFLATED, -rnd.randint(9, 15), rnd.randint(1, 9),
                             rnd.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        parts, o = [], 0
        while o < len(d):
            step = rnd.randint(1, max(1, len(d)))
            parts.append(c.compress(d[o:o + step]))
            o += step
            if rnd.random() < 0.3:
                parts.append(c.flush(rnd.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
        parts.append(c.flush())
        ms.append(synth.Member("r%d" % i, 8, b"".join(parts), len(d), zlib.crc32(d) & 0xFFFFFFFF, raw=d))
    img = synth.build_zip(ms)
    if huge_bytes:
        os.environ["OTZ_HUGE_BYTES"] = huge_bytes
    try:
        c = _ctx()
        _check(img, oracle, c)
        c.close()
    finally:
        os.environ.pop("OTZ_HUGE_BYTES", None)


@pytest.mark.parametrize("base", [0, 1, 7, 13])
def test_packed_output_arena_any_alignment(oracle, base):
    """The C-ABI puts no alignment requirement on out_ofs: entries packed back to back from an odd base offset —
    STORE, reference-container method 93, small DEFLATE, huge DEFLATE (segments executed in parallel: the symbol
    buffer follows the alignment of every segment's output) — must come out bit-exact."""
    rnd = random.Random(100 + base)
    ms = [synth.member("h0", synth.jsonlog_text((3 << 20) + 5, 41), 8),
          synth.member("st", synth.random_bytes(100003, 1), 0),
          synth.member("h1", synth.jsonlog_text((2 << 20) + 777, 42), 8, level=9),
          synth.member("zr", synth.jsonlog_text(200001, 43), 93),
          synth.member("h2", synth.jsonlog_text(1500001, 44), 8, level=1)]
    ms += [synth.member("s%d" % i, synth.jsonlog_text(rnd.randint(1, 90000), 50 + i), 8) for i in range(40)]
    img = synth.build_zip(ms)
    tab = parse_central(img).copy()
    pos = base
    for i in range(len(tab)):
        tab["out_ofs"][i] = pos
        pos += int(tab["uncomp_size"][i])
    c = _ctx()
    out, crc, st = c.extract_host(img, tab, default_opts())
    fb = int(c.L.otz_inflate_fallbacks(c.h))
    c.close()
    assert fb == 0, fb
    for i, m in enumerate(ms):
        assert (int(st[i]) & 0xFF) == 0 and not (int(st[i]) & 0x300), (m.name, hex(int(st[i])))
        o, n = int(tab["out_ofs"][i]), int(tab["uncomp_size"][i])
        assert bytes(out[o:o + n]) == m.raw, m.name
        assert int(crc[i]) == m.crc32


def test_parallel_segments_long_reach_and_short_segments(oracle):
    """Stress for the marker resolution: blocks of a few hundred symbols (memLevel 1-2: hundreds of segments per
    stream, most of them shorter than the 32 KiB window, more candidates than the 255 a stream keeps), matches at
    distances up to 32500 that cross many segment boundaries, byte runs (distance 1) running through block starts."""
    blob = synth.random_bytes(32500, 21)                                      # (zlib matches up to 32768 - 262 back)
    rnd = random.Random(22)
    d0 = blob * 70                                                            # every match reaches 32500 back
    d1 = b"".join(blob[rnd.randrange(30000):][:rnd.randint(3, 2000)] for _ in range(4000))   # far matches of all lengths
    d2 = b"".join(bytes([rnd.randrange(256)]) * rnd.randint(1, 70000) for _ in range(60))    # runs across block starts
    d3 = synth.jsonlog_text(2 << 20, 23)
    ms = []
    for i, d in enumerate((d0, d1, d2, d3)):
        for ml in (1, 2, 8):
            co = zlib.compressobj(9, zlib.DEFLATED, -15, ml)
            pl = co.compress(d) + co.flush()
            ms.append(synth.Member("p%d_%d" % (i, ml), 8, pl, len(d), zlib.crc32(d) & 0xFFFFFFFF, raw=d))
    img = synth.build_zip(ms)
    os.environ["OTZ_HUGE_BYTES"] = "200000"
    try:
        c = _ctx()
        fb, st, out = _check(img, oracle, c)
        c.close()
    finally:
        os.environ.pop("OTZ_HUGE_BYTES", None)
    assert fb <= 6, fb   # (streams that open with stored blocks — the first 32500 random bytes — are k_inflate's)


def test_stored_blocks_stay_on_the_fast_path(oracle):
    """Stored blocks WITH payload (dec:269-319) — incompressible data, level 0, stored blocks between dynamic ones, in
    small entries and inside the segments of huge ones: the tokenizer copies the payload into the literals of the
    stream (whole warp) instead of handing the stream to k_inflate."""
    rnd = random.Random(55)
    txt = synth.jsonlog_text(400000, 61)
    ms = [synth.member("rand", synth.random_bytes(200000, 62), 8),
          synth.member("lvl0", txt[:150000], 8, level=0),
          synth.member("mixed", txt[:90000] + synth.random_bytes(120000, 63) + txt[90000:200000], 8),
          synth.member("one", b"x", 8, level=0),
          synth.member("tiny0", synth.random_bytes(70, 64), 8, level=0),
          synth.member("huge_mixed", synth.jsonlog_text(2 << 20, 65) + synth.random_bytes(700000, 66) + synth.jsonlog_text(1 << 20, 67), 8),
          synth.member("huge_mixed2", synth.jsonlog_text(2 << 20, 65) + synth.random_bytes(700000, 66) + synth.jsonlog_text(100000, 67), 8),
          synth.member("huge_lvl0", synth.jsonlog_text(1500000, 68), 8, level=0),
          synth.member("huge_rand", synth.random_bytes(1300000, 69), 8)]
    for i in range(30):
        parts = [synth.jsonlog_text(rnd.randint(1, 40000), 70 + i) if rnd.random() < 0.5 else synth.random_bytes(rnd.randint(1, 90000), 100 + i)
                 for _ in range(rnd.randint(1, 5))]
        ms.append(synth.member("m%d" % i, b"".join(parts), 8, level=rnd.choice([0, 1, 6, 9])))
    img = synth.build_zip(ms)
    c = _ctx()
    fb, st, out = _check(img, oracle, c)
    c.close()
    # (a block that mixes random bytes and text can have a code with more long codes than the second-level tables of
    # the lane-per-stream decoder hold — such a stream is k_inflate's; without the stored-block path all of these were)
    assert fb <= 3, fb
