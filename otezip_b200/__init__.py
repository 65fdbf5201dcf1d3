"""otezip_b200 — B200-native archive-codec hot path of otezip (DEFLATE / method-93 / STORE + CRC-32).

The product is libotezip_b200.so (CUDA kernels + extern "C" shim + the plain-C libzip-subset
host library).  This package only holds the ctypes loader used by tests and bench.py and the
synthetic-archive generator; it contains no decoder and no CPU fallback.
"""
from .native import Lib, Ctx, OtzEntry, OtzOpts, lib_path  # noqa: F401
