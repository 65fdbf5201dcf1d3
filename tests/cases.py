"""Shared seeded test archives (inputs only)."""
import random
import zlib

from otezip_b200 import synth


def mixed_archive(seed: int = 7, n_tiny: int = 400, n_mid: int = 60, n_z: int = 20, n_s: int = 20):
    """Tiny entries (many violate the reference's end-of-block rule F1), every zlib strategy and
    level, full-flush points, stored blocks, method-93 containers with odd block sizes, STORE."""
    rnd = random.Random(seed)
    ms = []
    for i in range(n_tiny):
        n = rnd.randint(0, 300)
        d = synth.jsonlog_text(n, i) if i % 3 else bytes(rnd.randrange(256) for _ in range(n))
        ms.append(synth.member("t%d" % i, d, 8, ref_safe=False, level=rnd.choice([1, 6, 9])))
    for i in range(n_mid):
        n = rnd.randint(1000, 200000)
        d = synth.jsonlog_text(n, 1000 + i) if i % 5 else synth.random_bytes(n, i)
        strat = rnd.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])
        ms.append(synth.member("m%d" % i, d, 8, ref_safe=bool(i % 2), level=rnd.choice([0, 1, 6, 9]), strategy=strat,
                               full_flush_every=rnd.choice([0, 0, 4096, 65536])))
    for i in range(n_z):
        d = synth.jsonlog_text(rnd.randint(0, 300000), 2000 + i)
        ms.append(synth.member("z%d" % i, d, 93, block=rnd.choice([65535, 1000, 17]), block_type=rnd.choice([0, 2])))
    for i in range(n_s):
        d = synth.random_bytes(rnd.randint(0, 100000), i)
        ms.append(synth.member("s%d" % i, d, 0))
    rnd.shuffle(ms)
    return ms
