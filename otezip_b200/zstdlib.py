"""libzstd (system, 1.5.5 in this image) through ctypes: frame GENERATION for tests and bench only.
The product never calls it; the GPU decoder (k_zstd_tok.cuh) is checked against frames made here."""
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

    def compress_adv(self, data: bytes, level: int = 3, checksum: bool = False, content_size: bool = True, window_log: int = 0) -> bytes:
        """ZSTD_compress2 with frame parameters set: checksum flag, no Frame_Content_Size (then the frame carries a
        window descriptor instead of the single-segment flag), a small window (more, smaller matches)."""
        z = self.z
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_CCtx_setParameter.restype = C.c_size_t
        z.ZSTD_compress2.restype = C.c_size_t
        z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        cctx = z.ZSTD_createCCtx()
        try:
            for prm, val in ((100, level), (200, int(content_size)), (201, int(checksum))) + (((101, window_log),) if window_log else ()):
                assert not z.ZSTD_isError(z.ZSTD_CCtx_setParameter(cctx, prm, val))   # ZSTD_c_compressionLevel / contentSizeFlag / checksumFlag / windowLog
            cap = z.ZSTD_compressBound(len(data))
            buf = C.create_string_buffer(max(cap, 1))
            n = z.ZSTD_compress2(cctx, buf, cap, data, len(data))
            assert not z.ZSTD_isError(n)
            return buf.raw[:n]
        finally:
            z.ZSTD_freeCCtx(cctx)

    def decompress(self, frame: bytes, size: int) -> bytes:
        buf = C.create_string_buffer(max(size, 1))
        n = self.z.ZSTD_decompress(buf, size, frame, len(frame))
        assert not self.z.ZSTD_isError(n)
        return buf.raw[:n]
