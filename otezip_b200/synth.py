"""Synthetic ZIP32 archives for the BASELINE.json configurations.

Pure host-side data generation (Python zlib + struct): the JSON-log text model
and the on-disk records are the ones SURVEY.md Appendix A/D validated against
the reference (LFH/CDH/EOCD field layout: /root/reference/src/lib/otezip.c:
1443-1590; reader side :199-396).  Nothing here decodes anything.
"""
from __future__ import annotations

import random
import struct
import zlib
from dataclasses import dataclass

import numpy as np

M_STORE, M_DEFLATE, M_ZSTD = 0, 8, 93

_LEVELS = ("INFO", "WARN", "ERROR", "DEBUG")
_PATHS = ("/api/v1/items", "/api/v1/users", "/health", "/login", "/api/v2/orders/%d")
_STATUS = (200, 200, 200, 201, 404, 500)
_MSGS = ("ok", "not found", "timeout contacting upstream", "cache miss", "retrying")


def jsonlog_text(nbytes: int, seed: int) -> bytes:
    """JSON-log lines, truncated to nbytes."""
    rnd = random.Random(seed)
    ts = 1_700_000_000
    out, have = [], 0
    while have < nbytes:
        ts += rnd.randint(0, 3)
        p = _PATHS[rnd.randrange(5)]
        if "%d" in p:
            p = p % rnd.randint(1, 99999)
        line = '{"ts":%d,"level":"%s","path":"%s","status":%d,"latency_ms":%d,"user":"u%05d","msg":"%s"}\n' % (
            ts, _LEVELS[rnd.randrange(4)], p, _STATUS[rnd.randrange(6)], rnd.randint(1, 900),
            rnd.randint(0, 5000), _MSGS[rnd.randrange(5)])
        out.append(line)
        have += len(line)
    return "".join(out).encode()[:nbytes]


class TextPool:
    """A large JSON-log pool; entries are slices at seeded offsets (cheap bulk text)."""

    def __init__(self, pool_bytes: int = 64 << 20, seed: int = 1234):
        self.buf = jsonlog_text(pool_bytes, seed)
        self.rnd = random.Random(seed ^ 0x5EED)

    def take(self, nbytes: int) -> bytes:
        if nbytes >= len(self.buf):
            reps = nbytes // len(self.buf) + 1
            return (self.buf * reps)[:nbytes]
        o = self.rnd.randrange(0, len(self.buf) - nbytes + 1)
        return self.buf[o:o + nbytes]

    def offsets(self, sizes) -> list[int]:
        """The offsets take() would use for these sizes, in order, without cutting the slices (a rank that owns a
        shard of the entry list cuts only its own entries with at())."""
        return [-1 if n >= len(self.buf) else self.rnd.randrange(0, len(self.buf) - n + 1) for n in sizes]

    def at(self, offset: int, nbytes: int) -> bytes:
        if offset < 0:
            return (self.buf * (nbytes // len(self.buf) + 1))[:nbytes]
        return self.buf[offset:offset + nbytes]


def random_bytes(nbytes: int, seed: int) -> bytes:
    return np.random.default_rng(seed).bytes(nbytes)


def deflate_raw(data: bytes, level: int = 6, ref_safe: bool = True, strategy: int = zlib.Z_DEFAULT_STRATEGY,
                full_flush_every: int = 0) -> bytes:
    """Raw RFC 1951 stream.  ref_safe appends Z_SYNC_FLUSH before Z_FINISH so the
    reference's end-of-block rule always accepts it."""
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    parts = []
    if full_flush_every:
        for o in range(0, len(data), full_flush_every):
            parts.append(c.compress(data[o:o + full_flush_every]))
            if o + full_flush_every < len(data):
                parts.append(c.flush(zlib.Z_FULL_FLUSH))
    else:
        parts.append(c.compress(data))
    if ref_safe:
        parts.append(c.flush(zlib.Z_SYNC_FLUSH))
    parts.append(c.flush())
    return b"".join(parts)


def zstdref_container(data: bytes, block: int = 65535, block_type: int = 0, desc: int = 0x00) -> bytes:
    """The reference's method-93 raw-block container (SURVEY.md Appendix C;
    /root/reference/src/lib/zstd.inc.c:479-705): magic 28 B5 2F FD, one
    descriptor byte, blocks [last | type<<1][len lo][len hi] payload."""
    out = [b"\x28\xb5\x2f\xfd", bytes([desc])]
    n = len(data)
    if n == 0:
        out.append(bytes([1 | (0 << 1), 0, 0]))
        return b"".join(out)
    o = 0
    while o < n:
        ln = min(block, n - o)
        last = 1 if o + ln >= n else 0
        out.append(bytes([last | (block_type << 1), ln & 0xFF, ln >> 8]))
        out.append(data[o:o + ln])
        o += ln
    return b"".join(out)


@dataclass
class Member:
    name: str
    method: int
    payload: bytes          # bytes as stored in the archive
    uncomp_size: int
    crc32: int
    extra: bytes = b""      # LFH extra field (tests LFH name/extra skipping)
    raw: bytes | None = None  # original data when known (tests)


def member(name: str, data: bytes, method: int, **kw) -> Member:
    crc = zlib.crc32(data) & 0xFFFFFFFF
    if method == M_STORE:
        payload = data
    elif method == M_DEFLATE:
        payload = deflate_raw(data, **kw)
    elif method == M_ZSTD:
        payload = zstdref_container(data, **kw)
    else:
        raise ValueError(method)
    return Member(name, method, payload, len(data), crc, raw=data)


def build_zip(members: list[Member], comment: bytes = b"") -> bytes:
    """ZIP32 image with the field values the reference writer emits
    (otezip.c:1443-1590): needed 20, flags 0, made-by 0x031e, attr 0100644<<16."""
    out = bytearray()
    cd = bytearray()
    for m in members:
        name = m.name.encode()
        ofs = len(out)
        out += struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, m.method, 0, 0x21, m.crc32, len(m.payload),
                           m.uncomp_size, len(name), len(m.extra))
        out += name + m.extra + m.payload
        cd += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 0x031E, 20, 0, m.method, 0, 0x21, m.crc32,
                          len(m.payload), m.uncomp_size, len(name), 0, 0, 0, 0, 0o100644 << 16, ofs)
        cd += name
    cd_ofs = len(out)
    out += cd
    n = len(members)
    out += struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, n & 0xFFFF, n & 0xFFFF, len(cd), cd_ofs, len(comment))
    out += comment
    return bytes(out)


# ---------------------------------------------------------------- BASELINE configs

def config_c1(n: int = 1000, size: int = 65536, level: int = 6, seed: int = 1234) -> list[Member]:
    """configs[0]: n x 64 KiB JSON-log DEFLATE entries (the reference's CPU-runnable case)."""
    pool = TextPool(max(8 << 20, min(n * size, 64 << 20)), seed)
    return [member("log/%05d.json" % i, pool.take(size), M_DEFLATE, level=level) for i in range(n)]


def config_c3_sizes(n: int = 10000, seed: int = 3, lo: int = 12, hi: int = 24) -> list[int]:
    rnd = random.Random(seed)
    return [int(2 ** rnd.uniform(lo, hi)) for _ in range(n)]


def config_c3(n: int, seed: int = 3, lo: int = 12, hi: int = 24, level: int = 6) -> list[Member]:
    """configs[2]: mixed-size DEFLATE entries, sizes floor(2^U(lo,hi))."""
    pool = TextPool(64 << 20, seed)
    return [member("mix/%05d.log" % i, pool.take(s), M_DEFLATE, level=level)
            for i, s in enumerate(config_c3_sizes(n, seed, lo, hi))]


def config_c4(n: int, size: int = 262144, seed: int = 4) -> list[Member]:
    """configs[3] (4a): method-93 entries in the reference container."""
    pool = TextPool(max(8 << 20, min(n * size, 64 << 20)), seed)
    return [member("z/%05d.json" % i, pool.take(size), M_ZSTD) for i in range(n)]


def config_c2(n: int, size: int = 1 << 20, seed: int = 2) -> list[Member]:
    """configs[1]: STORE entries of seeded random bytes."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        d = rng.bytes(size)
        out.append(Member("r/%05d.bin" % i, M_STORE, d, size, zlib.crc32(d) & 0xFFFFFFFF, raw=d))
    return out
