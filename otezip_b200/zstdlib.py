"""libzstd (system, 1.5.5 in this image) through ctypes: frame GENERATION for tests and bench only.
The product never calls it; the GPU decoder (k_zstd.cuh) is checked against frames made here."""
from __future__ import annotations

import ctypes as C


class Zstd:
    def __init__(self):
        z = self.z = C.CDLL("libzstd.so.1")
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_versionNumber.restype = C.c_uint
        self.version = z.ZSTD_versionNumber()

    def compress(self, data: bytes, level: int = 3) -> bytes:
        cap = self.z.ZSTD_compressBound(len(data))
        buf = C.create_string_buffer(max(cap, 1))
        n = self.z.ZSTD_compress(buf, cap, data, len(data), level)
        assert not self.z.ZSTD_isError(n)
        return buf.raw[:n]

    def decompress(self, frame: bytes, size: int) -> bytes:
        buf = C.create_string_buffer(max(size, 1))
        n = self.z.ZSTD_decompress(buf, size, frame, len(frame))
        assert not self.z.ZSTD_isError(n)
        return buf.raw[:n]
