"""Round-2 additions on the GPU path: the multi-device host call, the pipelined host call, zip_open_from_source,
handles that outlive zip_close, directories that claim gigabytes, hostile chunk indexes, and the writer's headers
byte for byte against the reference's."""
import ctypes as C
import os
import resource
import struct
import subprocess
import time
import zlib

import numpy as np
import pytest

from otezip_b200 import Ctx, synth
from otezip_b200.native import Lib, OtzOpts, default_opts, parse_central
from otezip_b200.zipapi import ZipApi, ZipT
from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _multi(ctxs, img, tab, opts=None):
    """otz_extract_host_multi over the given contexts -> (out, crc, status)"""
    L = Lib.get().L
    L.otz_extract_host_multi.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(OtzOpts), C.c_void_p,
                                         C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    opts = opts or default_opts()
    n = len(tab)
    t = np.ascontiguousarray(tab)
    out_len = int((t["out_ofs"].astype(np.int64) + t["uncomp_size"]).max()) if n else 0
    out = np.zeros(max(out_len, 1), dtype=np.uint8)
    crc = np.zeros(max(n, 1), dtype=np.uint32)
    st = np.zeros(max(n, 1), dtype=np.int32)
    buf = np.frombuffer(img, dtype=np.uint8)
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = L.otz_extract_host_multi(arr, len(ctxs), buf.ctypes.data_as(C.c_void_p), buf.nbytes, t.ctypes.data_as(C.c_void_p), n, C.byref(opts),
                                  out.ctypes.data_as(C.c_void_p), out_len, crc.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), None)
    assert rc == 0, Lib.get().L.otz_last_error()
    return out[:out_len], crc[:n], st[:n]


@pytest.mark.parametrize("n_ctx", [2, 3])
def test_multi_device_call_equals_single_device_call(n_ctx):
    """otz_extract_host_multi: the table is cut by otz_partition, every context gets only its byte range
    (rebased rows) and runs on its own host thread.  Contexts may share a device, so one GPU is enough to check that
    the results — bytes, CRCs, status words, errors included — are those of the single-device call."""
    ms = cases.mixed_archive(seed=91, n_tiny=80, n_mid=50, n_z=6, n_s=6)
    ms += [synth.member("big%d" % i, synth.jsonlog_text((1 << 20) + 12345 * i, 60 + i), 8) for i in range(4)]
    img = bytearray(synth.build_zip(ms))
    tab = parse_central(bytes(img))
    for i in (5, 40, 90):   # corrupt a header and two payloads
        img[int(tab["lfh_ofs"][i]) + (0 if i == 5 else 30 + len(ms[i].name) + 3)] ^= 0x40
    img = bytes(img)
    one = Ctx(0)
    out1, crc1, st1 = one.extract_host(img, tab, default_opts())
    ctxs = [Ctx(0) for _ in range(n_ctx)]
    outm, crcm, stm = _multi(ctxs, img, tab)
    assert np.array_equal(st1, stm)
    for i in range(len(tab)):
        if (int(st1[i]) & 0xFF) == 0:
            o, n = int(tab["out_ofs"][i]), int(tab["uncomp_size"][i])
            assert np.array_equal(out1[o:o + n], outm[o:o + n]), i
            assert int(crc1[i]) == int(crcm[i]), i
    for c in ctxs + [one]:
        c.close()


@pytest.mark.parametrize("pipe_bytes", ["0", "300000", None])
def test_pipelined_host_call_equals_one_batch(pipe_bytes, oracle):
    """otz_extract_host cuts a batch into sub-batches that alternate on two lanes (H2D / kernels / D2H overlap).
    OTZ_PIPE_BYTES=0 turns the pipeline off, 300000 makes dozens of tiny sub-batches."""
    ms = cases.mixed_archive(seed=92, n_tiny=100, n_mid=60, n_z=8, n_s=8)
    img = synth.build_zip(ms)
    tab = parse_central(img)
    if pipe_bytes is not None:
        os.environ["OTZ_PIPE_BYTES"] = pipe_bytes
    try:
        c = Ctx(0)
        out, crc, st = c.extract_host(img, tab, default_opts())
        out2, crc2, st2 = c.extract_host(img, tab, default_opts())      # cached device buffers, second call
        c.close()
    finally:
        os.environ.pop("OTZ_PIPE_BYTES", None)
    assert np.array_equal(st, st2) and np.array_equal(crc, crc2) and np.array_equal(out, out2)
    rc, oents = oracle.load_central(img)
    ost, ocrc, oout, oofs = oracle.extract_all(img, oents)
    for i in range(len(tab)):
        ok = (int(st[i]) & 0xFF) == 0 and not (int(st[i]) & 0x300)
        assert ok == (ost[i] == 0), (i, hex(int(st[i])), int(ost[i]))
        if ok:
            n = int(tab["uncomp_size"][i])
            assert np.array_equal(out[int(tab["out_ofs"][i]):int(tab["out_ofs"][i]) + n], oout[int(oofs[i]):int(oofs[i]) + n]), i
            assert int(crc[i]) == int(ocrc[i])


def _tree(d):
    out = {}
    for root, _, fs in os.walk(d):
        for f in fs:
            q = os.path.join(root, f)
            out[os.path.relpath(q, d)] = open(q, "rb").read()
    return out


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "otezip_relinked")), reason="relinked CLI not built")
def test_cli_with_several_devices_extracts_the_same_tree(tmp_path):
    """OTEZIP_DEVICES shards every batch of the libzip read path over several contexts (here: the same GPU three times):
    the unchanged reference CLI, relinked, must extract what the reference CLI extracts."""
    z = os.path.join(G, "mixed.zip")
    outs = {}
    for tag, exe, env in (("ref", "otezip_ref", {}), ("multi", "otezip_relinked", {"OTEZIP_DEVICES": "0,0,0"}), ("all", "otezip_relinked", {"OTEZIP_DEVICES": "all"})):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([os.path.join(REFDIR, exe), "-x", z, "--verify-crc"], cwd=d, capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, **env))
        outs[tag] = (r.returncode, r.stdout, _tree(d))
    assert outs["ref"][0] == outs["multi"][0] == outs["all"][0] == 0
    assert outs["ref"][1] == outs["multi"][1] == outs["all"][1]
    assert outs["ref"][2] == outs["multi"][2] == outs["all"][2] and len(outs["multi"][2]) > 100


def test_zip_open_from_source(tmp_path):
    """otezip.c:1406-1440: an archive held in memory is opened through the temp-file trampoline."""
    api = ZipApi()
    L = api.L
    L.zip_open_from_source.restype = C.POINTER(ZipT)
    L.zip_open_from_source.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.zip_source_buffer_create.restype = C.c_void_p
    L.zip_source_buffer_create.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    files = [("a.json", synth.jsonlog_text(90000, 1)), ("b.bin", synth.random_bytes(5000, 2)), ("empty", b"")]
    img = synth.build_zip([synth.member(n, d, 8 if i == 0 else 0) for i, (n, d) in enumerate(files)])
    buf = C.create_string_buffer(img, len(img))
    src = L.zip_source_buffer_create(buf, len(img), 0, None)
    assert src
    za = L.zip_open_from_source(src, 0, None)
    assert za and L.zip_get_num_files(za) == 3
    for i, (n, d) in enumerate(files):
        assert L.zip_get_name(za, i, 0) == n.encode()
        zf = L.zip_fopen_index(za, i, 0)
        assert zf and zf.contents.size == len(d) and C.string_at(zf.contents.data, len(d)) == d
        L.zip_fclose(zf)
    assert L.zip_close(za) == 0
    L.zip_source_free(src)
    assert not L.zip_open_from_source(None, 0, None)
    junk = C.create_string_buffer(b"not a zip", 9)
    s2 = L.zip_source_buffer_create(junk, 9, 0, None)
    assert not L.zip_open_from_source(s2, 0, None)
    L.zip_source_free(s2)


def test_handle_outlives_zip_close(tmp_path):
    """In the reference every zip_file_t owns its buffer (otezip.c:1326-1331), so zf->data is still valid after zip_close;
    here handles point into a batch arena, which therefore has to stay until its last handle is closed."""
    api = ZipApi()
    L = api.L
    d0, d1 = synth.jsonlog_text(200000, 5), synth.jsonlog_text(3000, 6)
    p = tmp_path / "h.zip"
    p.write_bytes(synth.build_zip([synth.member("x", d0, 8), synth.member("y", d1, 8)]))
    err = C.c_int(0)
    za = L.zip_open(str(p).encode(), 0, C.byref(err))
    zf0, zf1 = L.zip_fopen_index(za, 0, 0), L.zip_fopen_index(za, 1, 0)
    assert zf0 and zf1
    assert L.zip_close(za) == 0
    junk = [bytearray(os.urandom(1 << 20)) for _ in range(8)]   # churn the heap
    assert C.string_at(zf0.contents.data, len(d0)) == d0
    buf = C.create_string_buffer(len(d1))
    assert L.zip_fread(zf1, buf, len(d1)) == len(d1) and buf.raw == d1
    assert L.zip_fclose(zf0) == 0 and L.zip_fclose(zf1) == 0
    del junk


def test_directory_claiming_gigabytes_allocates_nothing(tmp_path, capfd):
    """otezip.c:454-462 rejects an implausible entry BEFORE any allocation.  A 10 KB archive whose 64 entries each claim
    2 GiB must be answered the same way here: no pinned arena, no device scratch sized by the claim."""
    api = ZipApi()
    L = api.L
    real = b"A" * 40
    comp = zlib.compress(real, 9)[2:-4]
    ms = [synth.Member("bomb%d" % i, 8, comp, (2 << 30) - 1 - i, zlib.crc32(real)) for i in range(64)]
    ms.append(synth.member("fine", synth.jsonlog_text(5000, 1), 8))
    p = tmp_path / "bomb.zip"
    p.write_bytes(synth.build_zip(ms))
    rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    t0 = time.time()
    err = C.c_int(0)
    za = L.zip_open(str(p).encode(), 0, C.byref(err))
    assert za
    for i in range(64):
        assert not L.zip_fopen_index(za, i, 0)
    zf = L.zip_fopen_index(za, 64, 0)
    assert zf and C.string_at(zf.contents.data, zf.contents.size) == synth.jsonlog_text(5000, 1)
    L.zip_fclose(zf)
    assert L.zip_close(za) == 0
    assert time.time() - t0 < 20
    assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - rss0 < 600 * 1024      # KiB: far from 64 x 2 GiB
    assert capfd.readouterr().err.count("rejecting to avoid zipbomb") == 64


def _oz_extra(cb, csizes):
    body = struct.pack("<BBHII", 1, 0, 0, cb, len(csizes)) + b"".join(struct.pack("<I", c) for c in csizes)
    return struct.pack("<HH", 0x5A4F, len(body)) + body


def test_hostile_chunk_index_cannot_change_the_sequential_result(tmp_path, reflib, capfd):
    """a crafted 'OZ' chunk index must not make this library return other bytes than the reference's
    sequential decoder.  (1) chunk 0 ends with the stream's FINAL block: the sequential decoder stops there and
    zero-pads, so the chunk decode (which would happily continue with chunk 1) must be refused.  (2) chunk 0 ends on a
    Huffman block in the middle of a byte: the sequential decoder reads the next header from the bits left over."""
    api = ZipApi()
    part1, part2 = synth.jsonlog_text(65280, 1), synth.jsonlog_text(30000, 2)
    whole = part1 + part2
    # (1) two complete streams back to back, labelled as two chunks of one entry
    s1, s2 = synth.deflate_raw(part1, 6, False), synth.deflate_raw(part2, 6, False)
    m1 = synth.Member("final_inside", 8, s1 + s2, len(whole), zlib.crc32(whole), extra=_oz_extra(65280, [len(s1), len(s2)]))
    # (2) chunk 0 = a non-final Huffman block that ends mid-byte (zlib sync flush without its stored block: cut 4 bytes 00 00 FF FF
    # and the 3 header bits stay in the last byte), chunk 1 = a complete stream
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    a = co.compress(part1) + co.flush(zlib.Z_SYNC_FLUSH)
    assert a.endswith(b"\x00\x00\xff\xff")
    a = a[:-4]
    m2 = synth.Member("midbyte", 8, a + s2, len(whole), zlib.crc32(whole), extra=_oz_extra(65280, [len(a), len(s2)]))
    good = synth.member("good", whole, 8)
    p = tmp_path / "hostile.zip"
    p.write_bytes(synth.build_zip([m1, m2, good]))
    api.ref_compat.value = 1
    for verify in (0, 1):
        err, names, datas = api.read_all(str(p), verify_crc=verify)
        e2, ref = reflib.extract_file(str(p), verify_crc=verify)
        assert err == 0 and e2 == 0
        assert datas == ref, [None if d is None else len(d) for d in datas]
    capfd.readouterr()


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "otezip_relinked")), reason="relinked CLI not built")
def test_store_archive_headers_equal_the_reference_writer(tmp_path):
    """otezip.c:1443-1590: LFH / CDH / EOCD of a STORE-only archive written through this library are, field for field,
    what the reference writes — every byte except the DOS time and date (the two runs are seconds apart)."""
    files = {"a.txt": b"hello\n", "b.bin": synth.random_bytes(4096, 3), "dir name/c d.txt": b"x" * 1000, "empty": b""}
    src = tmp_path / "src"
    for n, d in files.items():
        q = src / n
        q.parent.mkdir(parents=True, exist_ok=True)
        q.write_bytes(d)
    blobs = {}
    for tag, exe in (("ref", "otezip_ref"), ("new", "otezip_relinked")):
        z = tmp_path / (tag + ".zip")
        r = subprocess.run([os.path.join(REFDIR, exe), "-c", str(z)] + list(files) + ["-z", "store"], cwd=src, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        blobs[tag] = bytearray(z.read_bytes())
    a, b = blobs["ref"], blobs["new"]
    assert len(a) == len(b)
    for blob in (a, b):   # blank the time / date fields: LFH +10..13, CDH +12..15
        pos = 0
        while True:
            pos = blob.find(b"PK\x03\x04", pos)
            if pos < 0:
                break
            blob[pos + 10:pos + 14] = b"\0\0\0\0"
            pos += 30
        pos = 0
        while True:
            pos = blob.find(b"PK\x01\x02", pos)
            if pos < 0:
                break
            blob[pos + 12:pos + 16] = b"\0\0\0\0"
            pos += 46
    assert a == b


def test_streaming_layout_with_data_descriptors(tmp_path, reflib):
    """SURVEY §8f rank 2, write side: with otezip_write_data_descriptors (or OTEZIP_DATA_DESCRIPTORS=1) zip_close emits
    the streaming layout — flag bit 3, zero CRC / sizes in the local header, a data descriptor behind the payload.
    Python's zipfile, the compiled reference (it reads CRC and sizes from the directory) and this library read it back."""
    import zipfile
    api = ZipApi()
    flag = C.c_int.in_dll(api.L, "otezip_write_data_descriptors")
    files = [("t.json", synth.jsonlog_text(200000, 7)), ("r.bin", synth.random_bytes(3000, 8)), ("empty", b""), ("s.txt", b"hello\n")]
    p = str(tmp_path / "dd.zip")
    flag.value = 1
    try:
        assert api.write_archive(p, files, 8) == 0
    finally:
        flag.value = 0
    raw = open(p, "rb").read()
    assert raw.count(b"PK\x07\x08") >= len(files)
    pos = 0
    for name, data in files:   # local headers: bit 3, zeros; descriptor: the real values
        assert raw[pos:pos + 4] == b"PK\x03\x04"
        flags, method, _, _, crc, comp, uncomp, nl, xl = struct.unpack_from("<HHHHIIIHH", raw, pos + 6)
        assert flags == 8 and crc == 0 and comp == 0 and uncomp == 0
        with zipfile.ZipFile(p) as z:
            info = z.getinfo(name)
        dpos = pos + 30 + nl + xl + info.compress_size
        assert struct.unpack_from("<IIII", raw, dpos) == (0x08074B50, zlib.crc32(data), info.compress_size, len(data))
        pos = dpos + 16
    with zipfile.ZipFile(p) as z:
        assert z.testzip() is None and [z.read(n) for n, _ in files] == [d for _, d in files]
        assert all(i.flag_bits & 8 for i in z.infolist())
    assert api.read_all(p)[2] == [d for _, d in files]
    e2, got = reflib.extract_file(p, verify_crc=1)
    assert e2 == 0 and got == [d for _, d in files]


def test_zstandard_writer_real_frames(tmp_path, capfd):
    """with otezip_zstd_frames (OTEZIP_ZSTD_FRAMES=1) `zip_set_file_compression(ZIP_CM_ZSTD)` produces REAL
    Zstandard frames (k_zstd_enc.cuh) instead of ending up at STORE.  Every payload is decoded by libzstd (the format's
    reference decoder) and by this library's own GPU decoder, bit for bit; incompressible and empty sources still fall
    back to STORE (otezip.c:793-801, :894-899); without the switch the reference's behaviour stays: everything STORE."""
    import zipfile
    from otezip_b200.zstdlib import Zstd
    api = ZipApi()
    zs = Zstd()
    rnd = __import__("random").Random(3)
    files = [("log.json", synth.jsonlog_text(300000, 11)), ("small.txt", b"hello hello hello hello hello hello\n"), ("empty", b""),
             ("rand.bin", synth.random_bytes(50000, 12)), ("runs", b"ab" * 40000 + b"x" * 70000), ("one", b"z"),
             ("far", synth.random_bytes(30000, 13) * 3), ("big.json", synth.jsonlog_text((1 << 20) + 777, 14)),
             ("mixed", synth.jsonlog_text(90000, 15) + synth.random_bytes(70000, 16) + synth.jsonlog_text(50000, 17)),
             ("lits", bytes(rnd.randrange(97, 123) for _ in range(20000)))]
    flag = C.c_int.in_dll(api.L, "otezip_zstd_frames")
    p = str(tmp_path / "z.zip")
    flag.value = 1
    try:
        assert api.write_archive(p, files, 93) == 0
        with zipfile.ZipFile(p) as z:
            infos = z.infolist()
        raw = open(p, "rb").read()
        n93 = 0
        for info, (name, data) in zip(infos, files):
            nl, xl = struct.unpack_from("<HH", raw, info.header_offset + 26)
            payload = raw[info.header_offset + 30 + nl + xl:info.header_offset + 30 + nl + xl + info.compress_size]
            assert info.file_size == len(data) and info.CRC == zlib.crc32(data)
            if info.compress_type == 93:
                n93 += 1
                assert payload[:4] == b"\x28\xb5\x2f\xfd" and len(payload) < len(data)
                assert zs.decompress(payload, len(data)) == data, name           # libzstd reads the frame
            else:
                assert info.compress_type == 0 and payload == data, name
        types = dict((i.filename, i.compress_type) for i in infos)
        print(types)
        assert n93 >= 4 and all(types[k] == 93 for k in ("log.json", "runs", "big.json", "mixed")), types
        assert all(types[k] == 0 for k in ("empty", "rand.bin", "one")), types   # not smaller than the input: STORE
        assert infos[0].compress_size * 4 < infos[0].file_size                     # the log text really is compressed
        err, names, datas = api.read_all(p)                                         # and so does the GPU decoder
        assert err == 0 and datas == [d for _, d in files]
    finally:
        flag.value = 0
    # the reference's observable behaviour without the switch: method 93 ends up at STORE, and real frames are not handed out
    q = str(tmp_path / "s.zip")
    assert api.write_archive(q, files[:2], 93) == 0
    with zipfile.ZipFile(q) as z:
        assert [i.compress_type for i in z.infolist()] == [0, 0]
    err, names, datas = api.read_all(p)
    assert datas[0] is None and datas[2] == b""
    capfd.readouterr()


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "otezip_relinked")), reason="relinked CLI not built")
def test_cli_create_with_zstd_frames(tmp_path):
    """`otezip -c a.zip files -z zstd` through the unchanged main.c (it sets za->default_method, main.c:188-191): with
    OTEZIP_ZSTD_FRAMES=1 the compressible file becomes a method-93 entry holding a frame libzstd decodes, and
    `otezip -x` (same switch) extracts the tree again."""
    import zipfile
    from otezip_b200.zstdlib import Zstd
    src = tmp_path / "src"
    src.mkdir()
    files = {"a.json": synth.jsonlog_text(150000, 21), "b.bin": synth.random_bytes(2000, 22)}
    for n, d in files.items():
        (src / n).write_bytes(d)
    z = tmp_path / "c.zip"
    env = dict(os.environ, OTEZIP_ZSTD_FRAMES="1")
    r = subprocess.run([os.path.join(REFDIR, "otezip_relinked"), "-c", str(z)] + list(files) + ["-z", "zstd"], cwd=src, capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0, r.stderr
    raw = z.read_bytes()
    with zipfile.ZipFile(z) as zf:
        infos = {i.filename: i for i in zf.infolist()}
    assert infos["a.json"].compress_type == 93 and infos["b.bin"].compress_type == 0
    i = infos["a.json"]
    nl, xl = struct.unpack_from("<HH", raw, i.header_offset + 26)
    payload = raw[i.header_offset + 30 + nl + xl:i.header_offset + 30 + nl + xl + i.compress_size]
    assert Zstd().decompress(payload, i.file_size) == files["a.json"]
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([os.path.join(REFDIR, "otezip_relinked"), "-x", str(z), "--verify-crc"], cwd=out, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and _tree(out) == files


def test_reverse_iteration_decodes_each_entry_once(tmp_path):
    """a request below every decoded window used to decode a full window forward from each index it touched
    (O(n x budget) for reverse iteration).  Windows now start up to half a budget earlier; reading 60 entries backwards
    with a 1 MiB budget must be right and must take a handful of batches, not 60."""
    api = ZipApi()
    L = api.L
    files = [("f%02d" % i, synth.jsonlog_text(40000 + 997 * i, 30 + i)) for i in range(60)]
    p = tmp_path / "rev.zip"
    p.write_bytes(synth.build_zip([synth.member(n, d, 8) for n, d in files]))
    os.environ["OTEZIP_BATCH_BYTES"] = str(1 << 20)
    try:
        err = C.c_int(0)
        za = L.zip_open(str(p).encode(), 0, C.byref(err))
        assert za
        ctxs = (C.c_void_p * 1)()
        launches0 = None
        for i in reversed(range(60)):
            zf = L.zip_fopen_index(za, i, 0)
            assert zf and C.string_at(zf.contents.data, zf.contents.size) == files[i][1], i
            L.zip_fclose(zf)
        # forward again, then a few random probes
        for i in (0, 59, 17, 18, 3, 40):
            zf = L.zip_fopen_index(za, i, 0)
            assert zf and C.string_at(zf.contents.data, zf.contents.size) == files[i][1], i
            L.zip_fclose(zf)
        assert L.zip_close(za) == 0
    finally:
        del os.environ["OTEZIP_BATCH_BYTES"]
