"""ZIP64 (beyond the reference, which stops at 65,535 entries and 4 GiB — F5): the directory
walk of the host library and of the Python mirror against Python's own zipfile module, an independent implementation."""
import ctypes as C
import io
import os
import zipfile
import zlib

import numpy as np
import pytest

from otezip_b200 import synth
from otezip_b200.native import parse_central
from otezip_b200.zipapi import ZipApi, ZipStat, ZIP_RDONLY

N = 70000


@pytest.fixture(scope="module")
def big_zip(tmp_path_factory):
    """70,000 small entries written by zipfile: > 65,535 entries forces the ZIP64 end-of-central-directory record."""
    p = tmp_path_factory.mktemp("z64") / "many.zip"
    with zipfile.ZipFile(p, "w", zipfile.ZIP_STORED, allowZip64=True) as z:
        for i in range(N):
            z.writestr("d/%05d.txt" % i, b"entry %d\n" % i if i % 7 else b"")
    return str(p)


def test_zip_open_walks_a_zip64_directory(big_zip):
    api = ZipApi()
    err = C.c_int(-99)
    za = api.L.zip_open(big_zip.encode(), ZIP_RDONLY, C.byref(err))
    assert za, err.value
    assert api.L.zip_get_num_files(za) == N
    st = ZipStat()
    for i in (0, 1, 65534, 65535, 65536, N - 1):
        assert api.L.zip_get_name(za, i, 0) == b"d/%05d.txt" % i
        api.L.zip_stat_init(C.byref(st))
        assert api.L.zip_stat_index(za, i, 0, C.byref(st)) == 0
        want = b"entry %d\n" % i if i % 7 else b""
        assert st.size == len(want) and st.crc == (zlib.crc32(want) & 0xFFFFFFFF)
    assert api.L.zip_name_locate(za, b"d/69999.txt", 0) == N - 1
    assert api.L.zip_close(za) == 0


def test_python_mirror_reads_the_same_table(big_zip):
    img = open(big_zip, "rb").read()
    tab = parse_central(img)
    assert len(tab) == N
    with zipfile.ZipFile(big_zip) as z:
        infos = z.infolist()
    for i in (0, 65535, 65536, N - 1):
        assert int(tab["lfh_ofs"][i]) == infos[i].header_offset and int(tab["crc32"][i]) == infos[i].CRC
        assert int(tab["uncomp_size"][i]) == infos[i].file_size


def test_zip64_escaped_offset_field():
    """A directory whose entries carry the 0xFFFFFFFF offset escape + extended information field (what a > 4 GiB archive
    has), built small by hand: both walkers must take the 64-bit offsets from the extra field."""
    import struct
    ms = [synth.member("a", b"hello\n", 0), synth.member("b", synth.jsonlog_text(5000, 1), 8)]
    out, cd = bytearray(), bytearray()
    for m in ms:
        name = m.name.encode()
        ofs = len(out)
        out += struct.pack("<IHHHHHIIIHH", 0x04034B50, 45, 0, m.method, 0, 0x21, m.crc32, len(m.payload), m.uncomp_size, len(name), 0)
        out += name + m.payload
        x = struct.pack("<HHQ", 1, 8, ofs)
        cd += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 0x031E, 45, 0, m.method, 0, 0x21, m.crc32, len(m.payload), m.uncomp_size,
                          len(name), len(x), 0, 0, 0, 0o100644 << 16, 0xFFFFFFFF) + name + x
    cd_ofs = len(out)
    out += cd
    rec = len(out)
    out += struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 0x031E, 45, 0, 0, len(ms), len(ms), len(cd), cd_ofs)
    out += struct.pack("<IIQI", 0x07064B50, 0, rec, 1)
    out += struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, 0xFFFF, 0xFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0)
    img = bytes(out)
    with zipfile.ZipFile(io.BytesIO(img)) as z:   # Python's reader agrees that this is a valid archive
        assert z.read("a") == b"hello\n" and z.testzip() is None
    tab = parse_central(img)
    assert [int(v) for v in tab["lfh_ofs"]] == [0, 30 + 1 + 6]
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".zip", delete=False) as f:
        f.write(img)
    try:
        api = ZipApi()
        err = C.c_int(-99)
        za = api.L.zip_open(f.name.encode(), ZIP_RDONLY, C.byref(err))
        assert za and api.L.zip_get_num_files(za) == 2
        rows = np.zeros(2, dtype=tab.dtype)
        api.L.otezip_b200_entry_table.restype = C.c_uint64
        api.L.otezip_b200_entry_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        assert api.L.otezip_b200_entry_table(za, rows.ctypes.data_as(C.c_void_p), 2) == 2
        assert [int(v) for v in rows["lfh_ofs"]] == [0, 37]
        assert api.L.zip_close(za) == 0
    finally:
        os.unlink(f.name)


@pytest.mark.gpu
def test_zip64_read_and_write_on_the_gpu_path(big_zip, tmp_path):
    """Every 997th entry of the 70,000-entry archive through zip_fopen_index; then a 66,000-entry archive written by
    zip_file_add / zip_close must open in Python's zipfile (ZIP64 record + locator) with every CRC intact."""
    api = ZipApi()
    err = C.c_int(-99)
    za = api.L.zip_open(big_zip.encode(), ZIP_RDONLY, C.byref(err))
    assert za
    for i in list(range(0, N, 997)) + [65535, 65536, N - 1]:
        zf = api.L.zip_fopen_index(za, i, 0)
        want = b"entry %d\n" % i if i % 7 else b""
        assert zf, i
        buf = C.create_string_buffer(64)
        got = api.L.zip_fread(zf, buf, 64)
        assert buf.raw[:got] == want
        api.L.zip_fclose(zf)
    assert api.L.zip_close(za) == 0
    out = tmp_path / "written64.zip"
    files = [("w/%05d" % i, (b"line %d\n" % i) * (1 + i % 5)) for i in range(66000)]
    assert api.write_archive(str(out), files, method=8) == 0
    with zipfile.ZipFile(out) as z:
        assert len(z.infolist()) == 66000
        assert z.testzip() is None
        assert z.read("w/65999") == files[65999][1]
