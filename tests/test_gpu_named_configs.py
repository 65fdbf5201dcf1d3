"""Per-entry byte parity at the BASELINE configurations' NAMED sizes (SURVEY.md §8d "Parity check"): every entry of
configs[0] (1,000 x 64 KiB), configs[3] (10,000 x 256 KiB, reference container) and configs[2] (10,000 entries,
4 KiB-16 MiB, 20.5 GB) is extracted by the GPU path through the C-ABI host call and compared — SHA-256 of the bytes,
size, accept/reject — with what the compiled reference (oracle/_ref) returns for the same archive, the reference
running on all host cores."""
import concurrent.futures as cf
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 4


def _ref_digests(reflib, img, tab, per_archive):
    """SHA-256 of every entry as the reference extracts it (zip_open / zip_fopen_index / zip_fclose, verify_crc = 1),
    the entries dealt to all cores longest first.  -> list of hex digests (None = zip_fopen_index returned NULL)"""
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    n = len(tab)
    paths = []
    for a in range(0, n, per_archive):
        lo = int(tab["lfh_ofs"][a])
        hi = int(tab["lfh_ofs"][a + per_archive]) if a + per_archive < n else len(img)
        blob = img[lo:hi].tobytes()
        p = os.path.join(tmp, "otz_named_%d_%d.zip" % (os.getpid(), a))
        with open(p, "wb") as f:
            f.write(blob[:blob.rfind(b"PK\x05\x06") + 22])
        paths.append(p)
    order = sorted(range(n), key=lambda i: -int(tab["uncomp_size"][i]))
    buckets = [order[t::NCPU] for t in range(NCPU)]
    out = [None] * n
    L = reflib.lib

    def work(idx):
        reflib.verify_crc.value = 1
        handles = {}
        for i in idx:
            f, k = divmod(i, per_archive)
            if f not in handles:
                err = C.c_int(0)
                handles[f] = L.zip_open(paths[f].encode(), 0, C.byref(err))
                assert handles[f], err.value
            zf = L.zip_fopen_index(handles[f], k, 0)
            if zf:
                sz = zf.contents.size
                out[i] = hashlib.sha256(C.string_at(zf.contents.data, sz) if sz else b"").hexdigest()
                L.zip_fclose(zf)
        for h in handles.values():
            L.zip_close(h)
    try:
        with cf.ThreadPoolExecutor(NCPU) as ex:
            list(ex.map(work, [b for b in buckets if b]))
    finally:
        for p in paths:
            os.unlink(p)
    return out


@pytest.mark.parametrize("name,per", [("c1", 60000), ("c4", 15000), ("c3", 60000)])
def test_named_config_every_entry_equals_the_reference(name, per, reflib):
    import bench
    from otezip_b200 import Ctx
    ctx = Ctx(0)
    wl = bench.workload(name, 0, 1, None, ctx.pinned)
    img, tab = wl["image"], wl["table"]
    n = len(tab)
    assert n == bench.DEFAULT_ENTRIES[name]
    out = ctx.pinned(wl["out_bytes"])
    crc = np.zeros(n, dtype=np.uint32)
    st = np.zeros(n, dtype=np.int32)
    ctx.lib.check(ctx.L.otz_extract_host(ctx.h, img.ctypes.data_as(C.c_void_p), img.nbytes, tab.ctypes.data_as(C.c_void_p), n, C.byref(wl["opts"]),
                                         out.ctypes.data_as(C.c_void_p), wl["out_bytes"], crc.ctypes.data_as(C.c_void_p),
                                         st.ctypes.data_as(C.c_void_p)), "otz_extract_host")
    ref = _ref_digests(reflib, img, tab, per)
    ofs, sz = tab["out_ofs"], tab["uncomp_size"]

    def dig(i):
        return hashlib.sha256(out[int(ofs[i]):int(ofs[i]) + int(sz[i])]).hexdigest()
    with cf.ThreadPoolExecutor(NCPU) as ex:
        mine = list(ex.map(dig, range(n)))
    accept = [bool(ctx.L.otz_status_accepts(int(s), 1, 1)) for s in st]
    bad = [i for i in range(n) if accept[i] != (ref[i] is not None) or (accept[i] and mine[i] != ref[i])]
    assert not bad, (len(bad), bad[:10])
    assert sum(accept) == n          # the generators emit streams the reference accepts
    assert np.array_equal(crc, tab["crc32"])
    ctx.pinned_free(out)
    ctx.pinned_free(img)
    ctx.close()
