"""The drop-in boundary: the reference's libzip-subset API served by libotezip_b200.so.  Scenarios follow
the reference's own tests (test/unit/test_empty_zip.c, test_set_file_compression.c, test/test.sh)."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import zipfile
import zlib

import pytest

from otezip_b200 import synth
from otezip_b200.zipapi import ZipApi, ZIP_CM_DEFLATE, ZIP_CM_STORE

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
EXPECTED = json.load(open(os.path.join(G, "expected.json")))
REFDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


@pytest.fixture(scope="module")
def api():
    return ZipApi()


def test_empty_zip(api, tmp_path):
    # test/unit/test_empty_zip.c:10-19
    p = tmp_path / "empty.zip"
    p.write_bytes(bytes([0x50, 0x4b, 0x05, 0x06] + [0] * 18))
    err, names, datas = api.read_all(str(p))
    assert err == 0 and names == [] and datas == []


def test_open_errors(api, tmp_path):
    err = C.c_int(0)
    assert not api.L.zip_open(str(tmp_path / "missing.zip").encode(), 0, C.byref(err)) and err.value == 11   # ZIP_ER_OPEN
    (tmp_path / "junk.zip").write_bytes(b"this is not a zip archive, not at all")
    assert not api.L.zip_open(str(tmp_path / "junk.zip").encode(), 0, C.byref(err)) and err.value == 21     # ZIP_ER_INCONS
    assert not api.L.zip_open(str(tmp_path / "x.zip").encode(), 1 | 2 | 8, C.byref(err)) and err.value == -1  # EXCL+TRUNCATE


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_read_path_matches_reference_golden(api, name):
    err, names, datas = api.read_all(os.path.join(G, name + ".zip"), verify_crc=1)
    assert err == 0
    exp = EXPECTED[name]
    assert len(datas) == len(exp)
    for i, e in enumerate(exp):
        if e is None:
            assert datas[i] is None, (name, i)
        else:
            d = datas[i]
            assert d is not None, (name, i)
            assert [len(d), zlib.crc32(d) & 0xFFFFFFFF, hashlib.sha256(d).hexdigest()] == e, (name, i)


def test_crc_warning_mode_returns_data(api, capfd):
    # otezip.c:674-677: without otezip_verify_crc a mismatch only warns
    err, names, datas = api.read_all(os.path.join(G, "edges.zip"), verify_crc=0)
    i = names.index(b"bad_crc")
    assert datas[i] is not None
    assert "Warning: CRC mismatch for 'bad_crc'" in capfd.readouterr().err


def test_set_file_compression_after_add(api, tmp_path, reflib):
    # test/unit/test_set_file_compression.c:10-127 (fails on the reference; must pass here)
    p = str(tmp_path / "c.zip")
    payload = b"A" * 4096
    assert api.write_archive(p, [("hello.txt", payload)], ZIP_CM_DEFLATE) == 0
    err, names, datas = api.read_all(p)
    assert names == [b"hello.txt"] and datas == [payload]
    with zipfile.ZipFile(p) as z:
        info = z.infolist()[0]
        assert info.compress_type == zipfile.ZIP_DEFLATED and info.compress_size < 200 and z.read("hello.txt") == payload
    e2, got = reflib.extract_file(p, verify_crc=1)           # the reference reads what we wrote
    assert e2 == 0 and got == [payload]


def test_write_path_mixed_methods_and_fallbacks(api, tmp_path, reflib):
    files = [("text.json", synth.jsonlog_text(300000, 1)), ("empty", b""), ("rand.bin", synth.random_bytes(70000, 2)),
             ("stored.txt", synth.jsonlog_text(5000, 3), ZIP_CM_STORE), ("tiny", b"hello\n"), ("name with spaces.txt", b"x" * 1000),
             ("zstd-falls-back-to-store", synth.jsonlog_text(4000, 4), 93)]
    p = str(tmp_path / "m.zip")
    assert api.write_archive(p, files, ZIP_CM_DEFLATE) == 0
    want = [f[1] for f in files]
    assert api.read_all(p)[2] == want
    e2, got = reflib.extract_file(p, verify_crc=1)
    assert e2 == 0 and got == want
    with zipfile.ZipFile(p) as z:
        infos = z.infolist()
        assert z.testzip() is None
        assert [i.compress_type for i in infos] == [8, 0, 0, 0, 0, 8, 0]     # fallbacks: empty, random, tiny, zstd -> STORE
        assert all(i.create_version == 30 and i.extract_version == 20 and i.external_attr == 0o100644 << 16 for i in infos)


def _tree(d):
    out = {}
    for root, _, fs in os.walk(d):
        for f in fs:
            q = os.path.join(root, f)
            out[os.path.relpath(q, d)] = open(q, "rb").read()
    return out


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "otezip_relinked")), reason="relinked CLI not built")
def test_unchanged_cli_relinked_against_the_gpu_library(tmp_path):
    """The reference's unmodified main.c linked against libotezip_b200.so (oracle/Makefile) must behave like
    the reference CLI: same stdout, same extracted files (main.c:429-585 extract, :177-262 create)."""
    ref, new = os.path.join(REFDIR, "otezip_ref"), os.path.join(REFDIR, "otezip_relinked")
    z = os.path.join(G, "mixed.zip")
    outs = {}
    for tag, exe in (("ref", ref), ("new", new)):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([exe, "-x", z, "--verify-crc"], cwd=d, capture_output=True, text=True, timeout=300)
        outs[tag] = (r.returncode, r.stdout, _tree(d))
    assert outs["ref"][0] == outs["new"][0] == 0
    assert outs["ref"][1] == outs["new"][1]
    assert outs["ref"][2] == outs["new"][2] and len(outs["new"][2]) > 100
    # create with the relinked CLI, extract with the reference CLI
    src = tmp_path / "src"
    src.mkdir()
    (src / "a.json").write_bytes(synth.jsonlog_text(200000, 5))
    (src / "b.bin").write_bytes(synth.random_bytes(30000, 6))
    (src / "c.txt").write_bytes(b"hello\n")
    r = subprocess.run([new, "-c", "out.zip", "a.json", "b.bin", "c.txt", "-z", "deflate"], cwd=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = tmp_path / "back"
    d.mkdir()
    r = subprocess.run([ref, "-x", str(src / "out.zip"), "--verify-crc"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0
    want = {k: v for k, v in _tree(src).items() if k != "out.zip"}
    assert _tree(d) == want
    assert os.path.getsize(src / "out.zip") < 80000     # a.json really got deflated


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "otezip_relinked")), reason="relinked CLI not built")
def test_cli_gzip_modes_on_the_gpu_path(tmp_path):
    """-g / -d of the unchanged CLI (main.c:590-832) through the zlib-named one-shot entry points (zcompat.c)."""
    import gzip
    new = os.path.join(REFDIR, "otezip_relinked")
    text = synth.jsonlog_text(400000, 9)
    rnd = synth.random_bytes(100000, 10)
    for name, data in (("t.json", text), ("r.bin", rnd), ("tiny", b"hello world\n")):
        (tmp_path / name).write_bytes(data)
        r = subprocess.run([new, "-g", name], cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        gz = (tmp_path / (name + ".gz")).read_bytes()
        assert gzip.decompress(gz) == data                           # CRC-32 + ISIZE trailer checked by Python's gzip
        if data is text:
            assert len(gz) < len(data) // 5
        # gunzip what gzip(1)-compatible tools produce, and what we produced
        (tmp_path / "std.gz").write_bytes(gzip.compress(data, 6))
        for src in ("std.gz", name + ".gz"):
            r = subprocess.run([new, "-d", src, "back.out"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr
            assert (tmp_path / "back.out").read_bytes() == data
            os.unlink(tmp_path / "back.out")


def _lfh_extra(img: bytes, lfh_ofs: int) -> bytes:
    import struct
    sig, = struct.unpack_from("<I", img, lfh_ofs)
    assert sig == 0x04034B50
    nl, xl = struct.unpack_from("<HH", img, lfh_ofs + 26)
    return img[lfh_ofs + 30 + nl:lfh_ofs + 30 + nl + xl]


def test_chunk_index_roundtrip_and_fallback(api, tmp_path, reflib):
    """Multi-chunk DEFLATE entries written by this library carry a chunk index in the LFH extra field; reading
    decodes the chunks in parallel.  The reference (which skips extra fields) reads the same archive, and a
    corrupted index must fall back to the sequential decode with identical bytes."""
    import struct
    files = [("big.json", synth.jsonlog_text(3 << 20, 31)), ("mid.json", synth.jsonlog_text(200000, 32)),
             ("one-chunk.json", synth.jsonlog_text(60000, 33)), ("rand.bin", synth.random_bytes(300000, 34)),
             ("mixed.bin", synth.jsonlog_text(150000, 35) + synth.random_bytes(140000, 36) + synth.jsonlog_text(100000, 37))]
    p = str(tmp_path / "idx.zip")
    assert api.write_archive(p, files, ZIP_CM_DEFLATE) == 0
    want = [f[1] for f in files]
    img = open(p, "rb").read()
    with zipfile.ZipFile(p) as z:
        assert z.testzip() is None
        infos = z.infolist()
    x = _lfh_extra(img, infos[0].header_offset)
    hid, hsz, ver = struct.unpack_from("<HHB", x, 0)
    cb, nc = struct.unpack_from("<II", x, 8)
    assert (hid, ver, cb) == (0x5A4F, 1, 65280) and nc == -(-len(want[0]) // 65280) and hsz == 12 + 4 * nc
    assert _lfh_extra(img, infos[2].header_offset) == b"" and _lfh_extra(img, infos[3].header_offset) == b""   # 1 chunk / STORE
    err, names, datas = api.read_all(p)
    assert datas == want
    e2, got = reflib.extract_file(p, verify_crc=1)          # the reference skips the extra field and decodes sequentially
    assert e2 == 0 and got == want
    # without the index: same bytes
    os.environ["OTEZIP_NO_INDEX"] = "1"
    try:
        assert api.read_all(p)[2] == want
    finally:
        del os.environ["OTEZIP_NO_INDEX"]
    # a lying index (two chunk sizes swapped: the sum still matches) must not change the result
    b = bytearray(img)
    xo = infos[0].header_offset + 30 + len(b"big.json") + 16
    s0, s1 = struct.unpack_from("<II", b, xo)
    assert s0 != s1
    struct.pack_into("<II", b, xo, s1, s0)
    q = str(tmp_path / "lie.zip")
    open(q, "wb").write(bytes(b))
    assert api.read_all(q)[2] == want


def test_append_and_replace(api, tmp_path, reflib):
    """Append mode (zip_open with ZIP_CREATE on an existing archive, otezip.c:730-733, :776-779) and zip_file_replace
    (otezip.c:1617-1663) on the batched writer: new entries and new bytes for an EXISTING entry are queued and written
    behind the old data by zip_close; the result must read back through this library and through the compiled
    reference."""
    from otezip_b200.zipapi import ZIP_CREATE
    import ctypes as C
    L = api.L
    p = tmp_path / "a.zip"
    f = [("f%d.json" % i, synth.jsonlog_text(50000 + 1000 * i, 300 + i)) for i in range(4)]
    assert api.write_archive(str(p), f, ZIP_CM_DEFLATE) == 0
    size0 = p.stat().st_size
    err = C.c_int(-99)
    za = L.zip_open(str(p).encode(), ZIP_CREATE, C.byref(err))
    assert za and err.value == 0 and L.zip_get_num_files(za) == 4

    def source(data):   # freep = 0: the caller keeps the buffer
        buf = C.create_string_buffer(data, max(len(data), 1))
        return L.zip_source_buffer(za, buf, len(data), 0), buf

    keep = []
    g0, g1 = synth.jsonlog_text(300000, 310), synth.random_bytes(20000, 311)
    s, b = source(g0); keep.append(b)
    i0 = L.zip_file_add(za, b"g0.json", s, 0)
    s, b = source(g1); keep.append(b)
    i1 = L.zip_file_add(za, b"g1.bin", s, 0)
    assert (i0, i1) == (4, 5)
    assert L.zip_set_file_compression(za, i0, ZIP_CM_DEFLATE, 0) == 0
    assert L.zip_set_file_compression(za, 2, ZIP_CM_STORE, 0) == -1          # on disk, untouched: cannot be relabelled
    new1, newg0 = synth.jsonlog_text(123457, 320), synth.jsonlog_text(77777, 321)
    s, b = source(new1); keep.append(b)
    assert L.zip_file_replace(za, 1, s, 0) == 0                               # an entry that is already in the archive
    L.zip_source_free(s)                                                      # (the reference leaves src to the caller)
    assert not L.zip_fopen_index(za, 1, 0)                                    # queued, not readable before zip_close
    s, b = source(newg0); keep.append(b)
    assert L.zip_file_replace(za, i0, s, 0) == 0                              # an entry added in this session
    L.zip_source_free(s)
    assert L.zip_file_replace(za, 99, s, 0) == -1
    zf = L.zip_fopen_index(za, 0, 0)                                          # existing entries stay readable
    assert zf and zf.contents.size == len(f[0][1])
    L.zip_fclose(zf)
    assert L.zip_close(za) == 0
    assert p.stat().st_size > size0
    want_names = [n.encode() for n, _ in f] + [b"g0.json", b"g1.bin"]
    want = [f[0][1], new1, f[2][1], f[3][1], newg0, g1]
    err2, names, datas = api.read_all(str(p))
    assert err2 == 0 and names == want_names and datas == want
    rerr, rdatas = reflib.extract_bytes(p.read_bytes(), verify_crc=1)
    assert rerr == 0 and rdatas == want
    # Python's zipfile agrees on the directory (the replaced entry's old bytes are dead space, not an entry)
    import zipfile
    with zipfile.ZipFile(str(p)) as z:
        assert z.testzip() is None and [i.filename.encode() for i in z.infolist()] == want_names
