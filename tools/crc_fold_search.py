#!/usr/bin/env python
"""Search for the lag set used by the fold CRC kernel (k_crc32.cuh).

CRC-32 is M(x) mod P.  If S(x) = x^D + sum_i x^(D - l_i) is a multiple of P, a bit of degree d >= D can be
replaced by bits l_i positions later in the stream:  u[t] = s[t] ^ XOR_i u[t - l_i].  With every lag of the
form l = R*m + r (R = 4096 bits = one warp row of 32 lanes x 16 bytes, |r| < 32) the update of a whole row is
a handful of 32-bit funnel shifts and XORs per lane — no table, no multiply.  This script finds such lag sets
(sum_i x^(-l_i) == 1 in GF(2)[x]/P) by meet in the middle and checks them by brute force against zlib.crc32.
"""
import itertools
import sys
import zlib

POLY = 0x104C11DB7
R = 4096


def mulmod(a, b):
    r = 0
    while b:
        if b & 1:
            r ^= a
        b >>= 1
        a <<= 1
        if a >> 32:
            a ^= POLY
    return r


def powmod(base, e):
    r = 1
    while e:
        if e & 1:
            r = mulmod(r, base)
        base = mulmod(base, base)
        e >>= 1
    return r


def search(M, k=4):
    xinv = powmod(2, (1 << 32) - 2)          # x^-1 (the multiplicative group has order 2^32-1)
    cands = []
    for m in range(1, M + 1):
        for r in range(-31, 32):
            if r < 0 and m < 2:
                continue
            cands.append((m, r, powmod(xinv, R * m + r)))
    pairs = {}
    for (i, a), (j, b) in itertools.combinations(enumerate(cands), 2):
        pairs.setdefault(a[2] ^ b[2], []).append((i, j))
    sols = []
    for v, lst in pairs.items():
        w = v ^ 1
        if w in pairs and v <= w:
            for (i, j) in lst:
                for (p, q) in pairs[w]:
                    if len({i, j, p, q}) == 4:
                        s = tuple(sorted({i, j, p, q}))
                        sols.append(s)
    sols = sorted(set(sols))
    out = []
    for s in sols:
        lags = [(cands[i][0], cands[i][1]) for i in s]
        cost = sum(1 for m, r in lags if r != 0)
        out.append((max(m for m, _ in lags), cost, lags))
    return sorted(out)


def check(lags, nbytes=20000, seed=1):
    """Bit-level model of the recurrence with K zero rows of padding; compare with zlib.crc32."""
    import random
    rnd = random.Random(seed)
    data = bytes(rnd.randrange(256) for _ in range(nbytes))
    mmax = max(m for m, _ in lags)
    K = mmax + 1
    bits = []
    for b in data:
        for i in range(8):
            bits.append((b >> i) & 1)
    n_real = (len(bits) + R - 1) // R * R
    pad_bits = n_real - len(bits)
    bits += [0] * (pad_bits + K * R)
    nrows = n_real // R
    u = bits[:]
    L = [R * m + r for m, r in lags]
    for t in range(len(u)):
        row = t // R
        for (m, r), l in zip(lags, L):
            src = t - l
            if src < 0:
                continue
            if src // R >= nrows:      # residue rows are never sources
                continue
            u[t] ^= u[src]
    resid = u[nrows * R:]
    # residue rows as a byte message; its pure remainder equals R(M || zeros(pad + K rows))
    rb = bytearray()
    for i in range(0, len(resid), 8):
        rb.append(sum(resid[i + j] << j for j in range(8)))

    def raw(msg):          # pure remainder: table algorithm with init 0, no final xor
        return zlib.crc32(msg, 0xFFFFFFFF) ^ 0xFFFFFFFF if False else crc_raw(msg)

    def crc_raw(msg):
        c = 0
        for b in msg:
            c ^= b
            for _ in range(8):
                c = (c >> 1) ^ (0xEDB88320 if c & 1 else 0)
        return c
    lhs = crc_raw(bytes(rb))
    rhs = crc_raw(data + b"\0" * ((pad_bits + K * R) // 8))
    return lhs == rhs


if __name__ == "__main__":
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    sols = search(M)
    print("M=%d: %d solutions" % (M, len(sols)))
    for mm, cost, lags in sols[:12]:
        print("max row lag %d, shifted terms %d, lags (rows, bit shift): %s  check=%s" % (mm, cost, lags, check(lags, 6000)))
