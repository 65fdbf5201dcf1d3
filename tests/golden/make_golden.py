"""Generates tests/golden/*.zip and expected.json by running the COMPILED, UNMODIFIED reference
(oracle/_ref/libotezip_ref.so, built from /root/reference by oracle/Makefile) over seeded
archives.  Run in the build container only:  python tests/golden/make_golden.py

expected.json: per archive, per entry: null when the reference's zip_fopen_index returned NULL
(otezip_verify_crc = 1), else [size, crc32, sha256 of the returned bytes].
"""
import hashlib
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import RefLib  # noqa: E402
from otezip_b200 import synth  # noqa: E402
from tests import cases  # noqa: E402


def archives():
    # (a) the reference tests' own known-answer payloads
    sent63 = b"Hello, this is a test of deflate compression and decompression."          # test_mzip_deflate.c:12
    sent60 = b"This is a test string for Zstandard compression in otezip!!"[:60]         # test_zstd.c:12 (60 bytes)
    pat10k = bytes(65 + (i % 26) for i in range(10000))                                   # test_zstd.c:82-95
    ka = [synth.member("hello.txt", b"A" * 4096, 8),                                      # test_set_file_compression.c:10-127
          synth.member("sentence.deflate", sent63, 8), synth.member("sentence.deflate.raw", sent63, 8, ref_safe=False),
          synth.member("sentence.zst", sent60, 93), synth.member("pattern.zst", pat10k, 93),
          synth.member("hello", b"hello\n", 0), synth.member("world", b"world\n", 8),     # test.sh:15-17,45-70
          synth.member("empty", b"", 0), synth.member("bytes256", bytes(range(256)), 8),  # test.sh:236-271
          synth.member("rand4k", synth.random_bytes(4096, 11), 8), synth.member("name with spaces.txt", b"spaces\n", 8),
          synth.member("dup", b"one\n", 0), synth.member("dup", b"two\n", 8),              # test.sh:288-317
          synth.member("helloworld", b"hello world\n", 8), synth.member("a5000", b"A" * 5000 + b"\n", 8),  # test-deflate.sh:16-22
          synth.member("rand10k", synth.random_bytes(10240, 12), 8)]
    yield "known_answers", synth.build_zip(ka)
    # (b) empty archive: the 22-byte EOCD of test_empty_zip.c:10-19
    yield "empty", synth.build_zip([])
    # (c) seeded mixed archive (tiny entries that trip F1, all strategies, method 93, STORE)
    yield "mixed", synth.build_zip(cases.mixed_archive(seed=7, n_tiny=120, n_mid=16, n_z=8, n_s=6))
    # (d) edge semantics: short / long streams, bad CRC, trailing bytes, zip-bomb ratio,
    #     bad method-93 magic (the reference writer's FD 2F B5 28), unknown block type, truncated stream
    d = synth.jsonlog_text(5000, 5)
    m_short = synth.member("short_stream", d, 8)
    m_short.uncomp_size = 6000                       # declared larger than produced: zero padded, CRC decides
    m_short2 = synth.member("short_stream_crc_ok", d, 8)
    m_short2.uncomp_size = 6000
    m_short2.crc32 = zlib.crc32(d + b"\0" * 1000) & 0xFFFFFFFF
    m_long = synth.member("long_stream", d, 8)
    m_long.uncomp_size = 4000                        # declared smaller: NULL
    m_crc = synth.member("bad_crc", d, 8)
    m_crc.crc32 ^= 1
    m_trail = synth.member("trailing_bytes", d, 8)
    m_trail.payload += b"\xde\xad\xbe\xef"
    m_bomb = synth.member("bomb", b"\0" * 3000000, 8, level=9)
    m_badmagic = synth.member("zstd_writer_magic", d, 93)
    m_badmagic.payload = b"\xfd\x2f\xb5\x28" + m_badmagic.payload[4:]
    m_rle = synth.member("zstd_rle_type", d, 93, block_type=1)
    m_trunc = synth.member("truncated", d, 8)
    m_trunc.payload = m_trunc.payload[:len(m_trunc.payload) // 2]
    m_store_sz = synth.member("store_size_mismatch", d, 0)
    m_store_sz.uncomp_size -= 1
    m_meth = synth.member("method_14", d, 0)
    m_meth.method = 14
    m_zl = synth.member("zstd_size_mismatch", d, 93)
    m_zl.uncomp_size += 1
    m_extra = synth.member("lfh_extra", d, 8)
    m_extra.extra = b"\x55\x54\x05\x00\x01\x00\x00\x00\x00"
    yield "edges", synth.build_zip([m_short, m_short2, m_long, m_crc, m_trail, m_bomb, m_badmagic, m_rle, m_trunc,
                                    m_store_sz, m_meth, m_zl, m_extra, synth.member("ok", d, 8)])


def main():
    ref = RefLib()
    exp = {}
    for name, img in archives():
        with open(os.path.join(HERE, name + ".zip"), "wb") as f:
            f.write(img)
        err, res = ref.extract_bytes(img, verify_crc=1)
        assert err == 0, (name, err)
        exp[name] = [None if r is None else [len(r), zlib.crc32(r) & 0xFFFFFFFF, hashlib.sha256(r).hexdigest()] for r in res]
        print(name, len(img), "bytes,", len(res), "entries,", sum(r is None for r in res), "rejected by the reference")
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(exp, f, indent=0)


if __name__ == "__main__":
    main()
