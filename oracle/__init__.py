"""ctypes bindings for the CPU oracle and the compiled reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by otezip_b200.

  Oracle   libotz_oracle.so        the plain-C restatement (oracle/otz_oracle.c)
  RefLib   _ref/libotezip_ref.so   the unmodified reference, built by oracle/Makefile
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libotz_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libotezip_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "otezip_ref")


def build(quiet: bool = True) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class OEntry(C.Structure):
    _fields_ = [("local_hdr_ofs", C.c_uint32), ("comp_size", C.c_uint32), ("uncomp_size", C.c_uint32),
                ("crc32", C.c_uint32), ("method", C.c_uint16), ("file_time", C.c_uint16),
                ("file_date", C.c_uint16), ("external_attr", C.c_uint32), ("name_ofs", C.c_uint32),
                ("name_len", C.c_uint16)]


class OOpts(C.Structure):
    _fields_ = [("verify_crc", C.c_int), ("ignore_zipbomb", C.c_int), ("max_ratio", C.c_uint64),
                ("max_slack", C.c_uint64)]


def default_opts(verify_crc: int = 1, ignore_zipbomb: int = 0) -> OOpts:
    return OOpts(verify_crc, ignore_zipbomb, 1000, 1 << 20)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.otzo_crc32.restype = C.c_uint32
        L.otzo_crc32.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        L.otzo_inflate_raw.restype = None
        L.otzo_inflate_raw.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.otzo_zstdref_decode.restype = C.c_int
        L.otzo_zstdref_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.otzo_load_central.restype = C.c_int
        L.otzo_load_central.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.POINTER(OEntry)), C.POINTER(C.c_uint32)]
        L.otzo_extract_range.restype = None
        L.otzo_extract_range.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(OEntry), C.c_uint32, C.c_uint32,
                                         C.POINTER(OOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def crc32(self, data: bytes, seed: int = 0) -> int:
        return self.lib.otzo_crc32(seed, data, len(data))

    def inflate_raw(self, comp: bytes, out_cap: int):
        """-> (ref_ret, rfc_ret, total_out, out_bytes[out_cap] zero-initialised like otezip.c:500)"""
        out = (C.c_uint8 * max(out_cap, 1))()
        tot, r1, r2 = C.c_uint32(), C.c_int(), C.c_int()
        self.lib.otzo_inflate_raw(comp, len(comp), out, out_cap, C.byref(tot), C.byref(r1), C.byref(r2))
        return r1.value, r2.value, tot.value, bytes(out)[:out_cap]

    def zstdref_decode(self, comp: bytes, out_cap: int):
        out = (C.c_uint8 * max(out_cap, 1))()
        tot = C.c_uint32()
        r = self.lib.otzo_zstdref_decode(comp, len(comp), out, out_cap, C.byref(tot))
        return r, tot.value, bytes(out)[:out_cap]

    def load_central(self, img: bytes):
        """-> (rc, list[OEntry])"""
        p = C.POINTER(OEntry)()
        n = C.c_uint32()
        rc = self.lib.otzo_load_central(img, len(img), C.byref(p), C.byref(n))
        ents = [OEntry.from_buffer_copy(p[i]) for i in range(n.value)] if rc == 0 else []
        if p:
            self.libc.free(p)
        return rc, ents

    def extract_all(self, img, ents: list[OEntry], opts: OOpts | None = None, first: int = 0, last: int | None = None):
        """Extract entries [first,last) -> (status int32[n], crc uint32[n], out uint8[], out_ofs uint64[n]).
        status: 0 accepted, 0x100 accepted with CRC warning, -1 zip_fopen_index would return NULL."""
        n = len(ents)
        last = n if last is None else last
        opts = opts or default_opts()
        arr = (OEntry * max(n, 1))(*ents)
        sizes = np.array([e.uncomp_size for e in ents], dtype=np.uint64)
        ofs = np.zeros(n, dtype=np.uint64)
        if n:
            ofs[1:] = np.cumsum(sizes)[:-1]
        out = np.zeros(int(sizes.sum()) + 1, dtype=np.uint8)
        crc = np.zeros(max(n, 1), dtype=np.uint32)
        st = np.full(max(n, 1), -2, dtype=np.int32)
        buf = img if isinstance(img, (bytes, bytearray)) else img.ctypes.data_as(C.c_void_p)
        self.lib.otzo_extract_range(buf, len(img), arr, first, last, C.byref(opts), out.ctypes.data_as(C.c_void_p),
                                    ofs.ctypes.data_as(C.c_void_p), crc.ctypes.data_as(C.c_void_p),
                                    st.ctypes.data_as(C.c_void_p))
        return st[:n], crc[:n], out, ofs


class ZipFileT(C.Structure):  # struct zip_file, src/include/otezip/zip.h:100-104
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("size", C.c_uint32), ("pos", C.c_uint64)]


class RefLib:
    """The compiled, unmodified reference (libzip-subset API, zip.h:192-215)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (oracle/_ref is built from /root/reference in the build container)")
        L = self.lib = C.CDLL(path)
        L.zip_open.restype = C.c_void_p
        L.zip_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.zip_close.argtypes = [C.c_void_p]
        L.zip_get_num_files.restype = C.c_uint64
        L.zip_get_num_files.argtypes = [C.c_void_p]
        L.zip_fopen_index.restype = C.POINTER(ZipFileT)
        L.zip_fopen_index.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.zip_fclose.argtypes = [C.POINTER(ZipFileT)]
        L.zip_get_name.restype = C.c_char_p
        L.zip_get_name.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        self.verify_crc = C.c_int.in_dll(L, "otezip_verify_crc")
        self.ignore_zipbomb = C.c_int.in_dll(L, "otezip_ignore_zipbomb")

    def open_bytes(self, img: bytes):
        f = tempfile.NamedTemporaryFile(prefix="otzref_", suffix=".zip", delete=False)
        f.write(img)
        f.close()
        return f.name

    def extract_file(self, path: str, verify_crc: int = 1, first: int = 0, last: int | None = None,
                     keep_data: bool = True):
        """zip_open -> zip_fopen_index(i) -> zip_fclose loop (the reference's own
        read path).  -> (err, [bytes|None per entry])"""
        self.verify_crc.value = verify_crc
        err = C.c_int(0)
        za = self.lib.zip_open(path.encode(), 0, C.byref(err))
        if not za:
            return err.value, None
        n = self.lib.zip_get_num_files(za)
        last = n if last is None else min(last, n)
        res = []
        for i in range(first, last):
            zf = self.lib.zip_fopen_index(za, i, 0)
            if not zf:
                res.append(None)
                continue
            if keep_data:
                res.append(C.string_at(zf.contents.data, zf.contents.size) if zf.contents.size else b"")
            else:
                res.append(zf.contents.size)
            self.lib.zip_fclose(zf)
        self.lib.zip_close(za)
        return 0, res

    def extract_indices(self, path: str, indices, verify_crc: int = 1, keep_data: bool = False):
        """zip_open -> zip_fopen_index(i) -> zip_fclose for the given entry numbers.  -> (err, [bytes | size | None])"""
        self.verify_crc.value = verify_crc
        err = C.c_int(0)
        za = self.lib.zip_open(path.encode(), 0, C.byref(err))
        if not za:
            return err.value, None
        res = []
        for i in indices:
            zf = self.lib.zip_fopen_index(za, i, 0)
            if not zf:
                res.append(None)
                continue
            if keep_data:
                res.append(C.string_at(zf.contents.data, zf.contents.size) if zf.contents.size else b"")
            else:
                res.append(zf.contents.size)
            self.lib.zip_fclose(zf)
        self.lib.zip_close(za)
        return 0, res

    def extract_bytes(self, img: bytes, verify_crc: int = 1):
        p = self.open_bytes(img)
        try:
            return self.extract_file(p, verify_crc)
        finally:
            os.unlink(p)
