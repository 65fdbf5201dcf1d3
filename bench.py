#!/usr/bin/env python
"""bench.py — extract GB/s (uncompressed output, device-timed) of the otezip_b200 hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4|c4z|c5|c3w|c2x] [--impl reference]

One "step" = one pass of the batched extract over one synthetic archive.  The default workload (the headline) is
BASELINE.json configs[2]: DEFLATE inflate of 10,000 mixed-size entries (4 KiB-16 MiB, JSON-log text, zlib level 6),
20.5 GB of output — the configuration the north star's target is quoted on; it fits one B200.  Under torchrun (N > 1)
the ONE archive is sharded: otz_partition cuts the entry table into N contiguous index ranges balanced by
comp + uncomp bytes, every rank owns one GPU and its byte range of the archive (strong scaling, no data-path
collective: torch.distributed only carries the barrier and the max / sum of the reported scalars); value = all
entries' bytes over the max-over-ranks device time.

Output: ONE JSON line (contract in the task statement) with `roofline` for the dominant kernels, `cpu_baseline` =
the compiled reference (oracle/_ref) on the host cores over a stride subsample of the same entry list, `e2e`
through the C-ABI host-buffer call, and — at N = 1 — `secondary`: the other BASELINE configurations (value,
roofline fraction, e2e) measured in the same process.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import struct
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GB = 1e9
NCPU = os.cpu_count() or 4

WORKLOADS = {
    "c3": "configs[2]: DEFLATE inflate of 10,000 mixed-size entries (4 KiB-16 MiB, synthetic log/JSON text)",
    "c1": "configs[0]: 1,000-entry DEFLATE archive, 64 KiB text-like entries, CRC-32 check",
    "c2": "configs[1]: STORE + CRC-32 verify only, 10,000 entries of 1 MiB random bytes",
    "c4": "configs[3]: method 93 decode of 10,000 entries of 256 KiB (the reference's method-93 container)",
    "c4z": "configs[3] shape with real Zstandard (RFC 8878) frames from libzstd level 3 (parity pinned by libzstd, not by the reference)",
    "c5": "configs[4]: archive creation, batched DEFLATE compress + CRC-32, 4 GiB synthetic corpus",
    "c5z": "configs[4] shape, method 93: batched Zstandard compress (real RFC 8878 frames) + CRC-32, 4 GiB synthetic corpus",
    "c5f": "configs[4] at compression level 1 (zip_set_file_compression flags 1-3: single-candidate parse), 4 GiB synthetic corpus",
    "c3w": "configs[2] shape, archive written by this library (chunk-indexed DEFLATE entries)",
    "c2x": "configs[1] shape, STORE entries extracted (copied to the arena) and CRC-checked",
}
DEFAULT_ENTRIES = {"c3": 10000, "c1": 1000, "c2": 10000, "c2x": 10000, "c4": 10000, "c4z": 10000, "c5": 16384, "c5z": 16384, "c5f": 16384, "c3w": 2000}
METRIC = {
    "c2": "CRC-32 verify GB/s (STORE payload bytes, device-timed)",
    "c5": "compress GB/s (uncompressed input, device-timed; CRC-32 + DEFLATE + compaction)",
    "c5z": "compress GB/s (uncompressed input, device-timed; CRC-32 + Zstandard + compaction)",
    "c5f": "compress GB/s (uncompressed input, device-timed; CRC-32 + DEFLATE level 1 + compaction)",
}
DECODE_KERNELS = {
    "c1": "k_inflate_spec + k_inflate_lz (+ k_inflate for declined streams)",
    "c3": "k_inflate_spec<1> + k_inflate_lz; entries >= 1 MiB: k_inflate_spec<4> + k_inflate_lz<symbols> + k_seg_window + k_seg_translate",
    "c3w": "k_inflate_spec + k_inflate_lz (+ k_inflate for declined streams)",
    "c4": "k_zstdref", "c4z": "k_zstd_lit + k_zstd_seq + k_inflate_lz<wide>", "c2": "k_store_copy", "c2x": "k_store_copy",
}
# stride of the entry subsample the CPU legs time (the reference needs ~1 min per pass of configs[2] on
# 16 cores; every stride-th entry of the same list keeps the size distribution)
C5_METHOD = {"c5": 8, "c5z": 93, "c5f": 8 | 0x100}   # 0x100 = OTZ_M_FAST (include/otz_gpu.h)
REF_STRIDE = {"c3": 16, "c1": 1, "c2": 16, "c2x": 16, "c4": 4, "c4z": 4, "c3w": 16, "c5": 16, "c5z": 16, "c5f": 16}


def metric_name(wl: str) -> str:
    return METRIC.get(wl, "extract GB/s (uncompressed output, device-timed)")


# ----------------------------------------------------------------------------- workloads
def entry_plan(name: str, n: int):
    """The deterministic part of a workload: sizes, method, text seed.  Same on every rank and in both arms."""
    from otezip_b200 import synth
    if name in ("c2", "c2x"):
        return dict(sizes=[1 << 20] * n, method=0, seed=2, sizes_desc="1 MiB each", data="seeded random bytes (PCG64)", codec="STORE")
    if name in ("c1",):
        return dict(sizes=[65536] * n, method=8, seed=1234, sizes_desc="64 KiB each", data="JSON-log text (synth.TextPool, seed 1234)",
                    codec="raw DEFLATE, zlib level 6, Z_SYNC_FLUSH + Z_FINISH tail (SURVEY F1)")
    if name in ("c3", "c3w"):
        return dict(sizes=synth.config_c3_sizes(n, seed=3, lo=12, hi=24), method=8, seed=3, sizes_desc="floor(2^U(12,24)), random.Random(3)",
                    data="JSON-log text (synth.TextPool, seed 3)", codec="raw DEFLATE, zlib level 6, Z_SYNC_FLUSH + Z_FINISH tail (SURVEY F1)")
    if name == "c4":
        return dict(sizes=[262144] * n, method=93, seed=4, sizes_desc="256 KiB each", data="JSON-log text (synth.TextPool, seed 4)",
                    codec="method 93, the reference's raw-block container (SURVEY F3)")
    if name == "c4z":
        return dict(sizes=[262144] * n, method=93, seed=4, sizes_desc="256 KiB each", data="JSON-log text (synth.TextPool, seed 4)",
                    codec="method 93, real Zstandard frames (libzstd level 3)")
    if name in ("c5", "c5z", "c5f"):
        return dict(sizes=[262144] * n, method=93 if name == "c5z" else 8, seed=5, sizes_desc="256 KiB each", data="JSON-log text (synth.TextPool, seed 5)",
                    codec=("DEFLATE" if name == "c5" else "DEFLATE level 1" if name == "c5f" else "Zstandard (real frames)") + " compress (GPU), zero-length / incompressible -> STORE")
    raise SystemExit("unknown workload " + name)


def wl_config(name: str, n: int, world: int, scaling: str) -> dict:
    """`config` of the JSON line: only what defines the workload — identical in the GPU arm and the reference arm."""
    pl = entry_plan(name, n)
    un = int(sum(pl["sizes"]))
    if world == 1:
        par = "1 GPU"
    elif scaling == "strong":
        par = "ONE archive sharded over %d GPUs: contiguous index ranges balanced by comp+uncomp bytes (otz_partition), no collective" % world
    else:
        par = "%d GPUs, one full archive each (weak scaling), no collective" % world
    return {"workload": WORKLOADS[name], "entries": n, "uncomp_bytes": un * (world if scaling == "weak" else 1), "entry_sizes": pl["sizes_desc"],
            "payload": pl["data"], "codec": pl["codec"], "parallelism": par,
            "cache": "inputs larger than L2 (no flush needed)" if un > 256e6 else "inputs smaller than L2: steady-state L2-resident"}


def _zip_headers(name: bytes, method: int, crc: int, comp: int, uncomp: int, lfh_ofs: int):
    lfh = struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, method, 0, 0x21, crc, comp, uncomp, len(name), 0) + name
    cdh = struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 0x031E, 20, 0, method, 0, 0x21, crc, comp, uncomp, len(name),
                      0, 0, 0, 0, 0o100644 << 16, lfh_ofs) + name
    return lfh, cdh


def build_archive_set(alloc, payloads_fn, indices, per_archive: int, method: int, threads: int = NCPU):
    """Lay out ceil(n/per_archive) ZIP32 archives back to back in one buffer from alloc(nbytes) (one ZIP32
    file holds at most 4 GiB / 65,535 entries).  indices = global entry numbers of this image;
    payloads_fn(i) -> (payload bytes-like, uncomp_size, crc32).  Returns (image, entry table, arena bytes)."""
    from otezip_b200.native import ENTRY_DTYPE
    indices = list(indices)
    with cf.ThreadPoolExecutor(threads) as ex:
        metas = list(ex.map(payloads_fn, indices))
    n = len(indices)
    total = 0
    layout = []
    for a in range(0, n, per_archive):
        base = total
        pos = cd_len = 0
        recs = []
        for k in range(a, min(n, a + per_archive)):
            name = b"e/%05d.bin" % indices[k]
            recs.append((k, pos, name))
            pos += 30 + len(name) + len(metas[k][0])
            cd_len += 46 + len(name)
        layout.append((base, recs, pos, cd_len))
        total = (total + pos + cd_len + 22 + 15) & ~15
    img = alloc(total + 64)
    tab = np.zeros(n, dtype=ENTRY_DTYPE)
    out_ofs = 0
    for base, recs, cd_ofs, cd_len in layout:
        cd = bytearray()
        for k, pos, name in recs:
            payload, uncomp, crc = metas[k]
            lfh, cdh = _zip_headers(name, method, crc, len(payload), uncomp, pos)
            o = base + pos
            img[o:o + len(lfh)] = np.frombuffer(lfh, dtype=np.uint8)
            o += len(lfh)
            img[o:o + len(payload)] = np.frombuffer(payload, dtype=np.uint8)
            cd += cdh
            tab[k] = (base + pos, out_ofs, len(payload), uncomp, crc, method, 0)
            out_ofs += (uncomp + 15) & ~15
            metas[k] = None
        o = base + cd_ofs
        img[o:o + len(cd)] = np.frombuffer(bytes(cd), dtype=np.uint8)
        o += len(cd)
        eocd = struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, len(recs) & 0xFFFF, len(recs) & 0xFFFF, len(cd), cd_ofs, 0)
        img[o:o + 22] = np.frombuffer(eocd, dtype=np.uint8)
    return img[:total], tab, out_ofs


def shard_of(name: str, sizes, world: int, rank: int):
    """This rank's contiguous index range of the ONE archive (strong scaling).  The split is otz_partition over the
    directory sizes; so that a rank only has to compress its own shard, the compressed sizes that enter the weights
    are the size model of this text (ratio 9.3 for DEFLATE level 6, 7.6 for zstd-3, 1 for STORE / the container)."""
    from otezip_b200.native import ENTRY_DTYPE, partition
    if world == 1:
        return 0, len(sizes)
    ratio = {"c1": 9.3, "c3": 9.3, "c3w": 6.9, "c4z": 7.6}.get(name, 1.0)
    t = np.zeros(len(sizes), dtype=ENTRY_DTYPE)
    t["uncomp_size"] = np.array(sizes, dtype=np.uint32)
    t["comp_size"] = (np.array(sizes, dtype=np.float64) / ratio).astype(np.uint32)
    first = partition(t, world)
    return int(first[rank]), int(first[rank + 1])


def workload(name: str, rank: int, world: int, n_entries: int | None, alloc, scaling: str = "strong", indices=None):
    """-> dict(image, table, out_bytes, uncomp_bytes, algo_bytes, opts, dominant, n_entries, lo, hi).
    indices: explicit global entry numbers (the CPU legs' stride subsample) instead of this rank's shard."""
    from otezip_b200 import synth
    from otezip_b200.native import default_opts
    if name == "c3w":
        return workload_c3w(rank, n_entries, alloc)
    n = n_entries or DEFAULT_ENTRIES[name]
    pl = entry_plan(name, n)
    sizes, method = pl["sizes"], pl["method"]
    seed_shift = 7919 * rank if (scaling == "weak" and world > 1) else 0
    if indices is None:
        lo, hi = shard_of(name, sizes, world, rank) if scaling == "strong" else (0, n)
        indices = range(lo, hi)
    else:
        lo, hi = 0, n
    if name in ("c2", "c2x"):
        def gen(i):
            rng = np.random.Generator(np.random.PCG64([2, seed_shift, i]))
            d = rng.bit_generator.random_raw(sizes[i] // 8).view(np.uint8)
            return d, sizes[i], zlib.crc32(d) & 0xFFFFFFFF
        per = 1250
    else:
        pool = synth.TextPool(64 << 20, seed=pl["seed"] + seed_shift)
        offs = pool.offsets(sizes)
        zs = None
        if name == "c4z":
            from otezip_b200.zstdlib import Zstd
            zs = Zstd()

        def gen(i):
            d = pool.at(offs[i], sizes[i])
            crc = zlib.crc32(d) & 0xFFFFFFFF
            pay = synth.deflate_raw(d, 6, True) if method == 8 else (zs.compress(d, 3) if zs else synth.zstdref_container(d))
            return pay, len(d), crc
        per = 15000 if name == "c4" else 60000
    img, tab, out_bytes = build_archive_set(alloc, gen, indices, per, method)
    un = int(tab["uncomp_size"].astype(np.int64).sum())
    comp = int(tab["comp_size"].astype(np.int64).sum())
    if name == "c2":
        return dict(image=img, table=tab, out_bytes=0, uncomp_bytes=un, algo_bytes=un, opts=default_opts(verify_only=1), dominant="crc",
                    n_entries=n, lo=lo, hi=hi)
    return dict(image=img, table=tab, out_bytes=out_bytes, uncomp_bytes=un, algo_bytes=(2 * un if name == "c2x" else un + comp),
                opts=default_opts(), dominant="decode", n_entries=n, lo=lo, hi=hi)


def workload_c3w(rank: int, n_entries: int | None, alloc):
    """configs[2] shape (mixed 4 KiB-16 MiB JSON-log entries), but the archive is WRITTEN BY THIS LIBRARY's GPU
    compressor: multi-chunk entries carry the chunk index, so large entries decode chunk-parallel."""
    from otezip_b200 import Ctx, synth
    from otezip_b200.native import default_opts, expand_chunk_index, ENTRY_DTYPE
    n = n_entries or 2000
    sizes = synth.config_c3_sizes(n, seed=3 + rank, lo=12, hi=24)
    pool = synth.TextPool(64 << 20, seed=3 + 7919 * rank)
    ctx = Ctx(int(os.environ.get("LOCAL_RANK", "0")))
    in_ofs = np.zeros(n, dtype=np.uint64)
    pos = 0
    for i, sz in enumerate(sizes):
        in_ofs[i] = pos
        pos += (sz + 15) & ~15
    src = ctx.pinned(pos)
    crcs = []
    for i, sz in enumerate(sizes):
        d = pool.take(sz)
        src[int(in_ofs[i]):int(in_ofs[i]) + sz] = np.frombuffer(d, dtype=np.uint8)
        crcs.append(zlib.crc32(d) & 0xFFFFFFFF)
    in_len = np.array(sizes, dtype=np.uint32)
    d_in = ctx.dev_alloc(pos)
    ctx.h2d(d_in, src)
    job = ctx.deflate_plan(in_ofs, in_len, np.full(n, 8, dtype=np.uint16))
    ctx.deflate_run(job, d_in, pos)
    ofs, csz, crc, meth, total = ctx.deflate_results(job, n)
    assert [int(c) for c in crc] == crcs
    comp = ctx.deflate_fetch(job, total)
    first, cnt, cs, cb = ctx.deflate_chunks(job, n)
    ctx.deflate_destroy(job)
    ctx.dev_free(d_in)
    ctx.pinned_free(src)
    ctx.close()
    # ZIP32 image with the chunk index in the LFH extra field (what otezip.c:finalize_archive writes)
    parts, tabrows = [], []
    o = 0
    out_ofs = 0
    for i in range(n):
        name = b"w/%05d.log" % i
        extra = b""
        if meth[i] == 8 and 2 <= cnt[i] <= 16000:
            body = struct.pack("<BBHII", 1, 0, 0, cb, int(cnt[i])) + cs[int(first[i]):int(first[i]) + int(cnt[i])].astype("<u4").tobytes()
            extra = struct.pack("<HH", 0x5A4F, len(body)) + body
        payload = comp[int(ofs[i]):int(ofs[i]) + int(csz[i])]
        lfh = struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, int(meth[i]), 0, 0x21, crcs[i], int(csz[i]), sizes[i], len(name), len(extra))
        tabrows.append((o, out_ofs, int(csz[i]), sizes[i], crcs[i], int(meth[i]), 0))
        parts += [lfh, name, extra, payload.tobytes()]
        o += len(lfh) + len(name) + len(extra) + int(csz[i])
        out_ofs += (sizes[i] + 15) & ~15
    blob = b"".join(parts)
    img = alloc(len(blob) + 64)
    img[:len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    tab = expand_chunk_index(img, np.array(tabrows, dtype=ENTRY_DTYPE))
    un = int(sum(sizes))
    return dict(image=img, table=tab, out_bytes=out_ofs, uncomp_bytes=un, algo_bytes=un + int(total), opts=default_opts(),
                dominant="decode", n_entries=n, lo=0, hi=n)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, pci_bus_id: str, period: float = 0.02):
        super().__init__(daemon=True)
        self.samples = []
        self.period = period
        self.stop_ev = threading.Event()
        self.ok = False
        self.marks = []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode())
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_ev.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        self.stop_ev.set()
        self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        t0, t1 = (self.marks + [None, None])[:2]
        sel = [s for s in self.samples if t0 is not None and t1 is not None and t0 <= s[0] <= t1]
        where = "timed region"
        if len(sel) < 2:
            sel, where = self.samples, "whole bench (timed region shorter than the sampling period)"
        sm = sorted(s[1] for s in sel)
        bits = 0
        for s in sel:
            bits |= s[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm,
                "reasons": [v for k, v in self.REASONS.items() if bits & k], "samples": len(sel), "window": where}


# ----------------------------------------------------------------------------- distributed plumbing
class Dist:
    def __init__(self, n):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local)
            dist.init_process_group(backend)
            self.dev = torch.device("cuda", self.local) if backend == "nccl" else torch.device("cpu")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
            if self.dev.type == "cuda":
                self.torch.cuda.synchronize()

    def _red(self, v: float, op) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v: float) -> float:
        return self._red(v, self.dist.ReduceOp.MAX) if self.world > 1 else v

    def sum(self, v: float) -> float:
        return self._red(v, self.dist.ReduceOp.SUM) if self.world > 1 else v

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------- the reference on the host cores
def numa_cpus_of_gpu(pci_bus_id: str):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None."""
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % pci_bus_id.lower()).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return None


def reference_sample(wl_name: str, n_entries: int | None, tmpdir: str):
    """The CPU legs' input: every stride-th entry of the workload's own entry list (same sizes, same text, same codec)
    as ZIP32 file(s) on disk (the reference reads FILE*).  -> (paths, per-file entry sizes, bytes, description)"""
    n = n_entries or DEFAULT_ENTRIES[wl_name]
    name = "c4" if wl_name == "c4z" else ("c3" if wl_name in ("c3w", "c5", "c5z", "c5f") else wl_name)   # the reference rejects real Zstandard frames (F3)
    stride = REF_STRIDE[wl_name]
    idx = list(range(0, n, stride))
    bufs = []
    wl = workload(name, 0, 1, n, lambda nbytes: bufs.append(np.zeros(nbytes, dtype=np.uint8)) or bufs[-1], indices=idx)
    img, tab = wl["image"], wl["table"]
    # build_archive_set lays ZIP32 archives back to back: one file per archive
    paths, sizes = [], []
    per = 1250 if name in ("c2", "c2x") else 15000 if name == "c4" else 60000
    for a in range(0, len(idx), per):
        t = tab[a:a + per]
        lo = int(t["lfh_ofs"][0])
        hi = int(tab["lfh_ofs"][a + per]) if a + per < len(idx) else len(img)
        p = os.path.join(tmpdir, "otz_ref_sample_%s_%d_%d.zip" % (wl_name, os.getpid(), a))
        blob = img[lo:hi].tobytes()
        with open(p, "wb") as f:
            f.write(blob[:blob.rfind(b"PK\x05\x06") + 22])   # (without the alignment padding behind the EOCD record)
        paths.append(p)
        sizes.append(t["uncomp_size"].astype(np.int64))
    what = "every %d%s entry of the workload's entry list (%d entries, %.3f GB)" % (
        stride, "th" if stride != 1 else "st", len(idx), wl["uncomp_bytes"] / GB)
    if wl_name == "c4z":
        what += "; reference-container payloads — the reference rejects real Zstandard frames (SURVEY F3)"
    if wl_name in ("c5", "c5z", "c5f"):
        what += "; the reference's DEFLATE writer is broken (SURVEY F2), so the CPU leg is its READ path over the same text"
    return paths, sizes, wl["uncomp_bytes"], what


def reference_pass(ref, paths, sizes, threads: int) -> float:
    """zip_open -> zip_fopen_index(i) -> zip_fclose (otezip_verify_crc = 1) over every entry of the sample, the entries
    dealt to `threads` workers longest first (the reference is single-threaded: one archive handle per worker).
    Returns seconds."""
    jobs = sorted(((int(s), f, i) for f, sz in enumerate(sizes) for i, s in enumerate(sz)), reverse=True)
    buckets = [[] for _ in range(threads)]
    load = [0] * threads
    for s, f, i in jobs:   # LPT
        t = load.index(min(load))
        buckets[t].append((f, i))
        load[t] += s + 4096

    def one(b):
        byfile = {}
        for f, i in b:
            byfile.setdefault(f, []).append(i)
        for f, ii in byfile.items():
            err, res = ref.extract_indices(paths[f], ii, verify_crc=1)
            assert err == 0 and all(r is not None for r in res)
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, [b for b in buckets if b]))
    return time.perf_counter() - t0


def cpu_baseline(wl_name: str, n_entries: int | None):
    from oracle import RefLib
    ref = RefLib()
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    paths, sizes, nbytes, what = reference_sample(wl_name, n_entries, tmp)
    try:
        reference_pass(ref, paths, sizes, NCPU) if nbytes < 300e6 else None   # warm the page cache for small samples
        dt = reference_pass(ref, paths, sizes, NCPU)
    finally:
        for p in paths:
            os.unlink(p)
    return {"value": nbytes / dt / GB, "unit": "GB/s", "cores": NCPU, "kind": "reference",
            "sample": "%s, one pass on %d threads (compiled reference oracle/_ref: zip_open / zip_fopen_index / zip_fclose, "
                      "otezip_verify_crc=1)" % (what, NCPU)}


def run_reference(args, dist: Dist):
    if dist.rank != 0:
        return
    from oracle import RefLib
    ref = RefLib()
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    paths, sizes, nbytes, what = reference_sample(args.workload, args.entries, tmp)
    try:
        for _ in range(args.warmup):
            reference_pass(ref, paths, sizes, NCPU)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            reference_pass(ref, paths, sizes, NCPU)
        dt = time.perf_counter() - t0
    finally:
        for p in paths:
            os.unlink(p)
    v = args.steps * nbytes / dt / GB
    n = args.entries or DEFAULT_ENTRIES[args.workload]
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.workload),
        "timing": "host wall clock: the reference's own CPU implementation (compiled from its sources, oracle/_ref) on all host cores",
        "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": wl_config(args.workload, n, max(1, args.gpus), args.scaling),
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": NCPU, "kind": "reference",
                         "sample": "each step: " + what + ", %d threads" % NCPU},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------- GPU arm: the write path (configs[4])
def run_c5(ctx, dist: "Dist", n_entries, steps, warmup, e2e_steps, sampler=None, method=8):
    """configs[4]: archive creation — batched DEFLATE compress + CRC-32 of a synthetic corpus (default 16,384 x 256 KiB
    JSON-log = 4 GiB), device-timed; every stream verified through zlib and the compiled reference."""
    import ctypes as C
    from otezip_b200 import synth
    n = n_entries or DEFAULT_ENTRIES["c5"]
    size = 262144
    lo, hi = (0, n) if dist.world == 1 else shard_of("c5", [size] * n, dist.world, dist.rank)
    m = hi - lo
    pool = synth.TextPool(64 << 20, seed=5)
    offs = pool.offsets([size] * n)
    img = ctx.pinned(max(m, 1) * size)
    srcs = []
    for k, i in enumerate(range(lo, hi)):
        d = pool.at(offs[i], size)
        img[k * size:(k + 1) * size] = np.frombuffer(d, dtype=np.uint8)
        srcs.append(d)
    in_ofs = np.arange(m, dtype=np.uint64) * size
    in_len = np.full(m, size, dtype=np.uint32)
    meth = np.full(m, method, dtype=np.uint16)
    d_in = ctx.dev_alloc(img.nbytes)
    ctx.h2d(d_in, img)
    job = ctx.deflate_plan(in_ofs, in_len, meth)
    ctx.sync()
    for _ in range(warmup):
        ctx.deflate_run(job, d_in, img.nbytes)
    ctx.sync()
    dist.barrier()
    ctx.sync()
    l0 = ctx.launches()
    if sampler:
        sampler.mark()
    ctx.timer_start()
    for _ in range(steps):
        ctx.deflate_run(job, d_in, img.nbytes)
    ms = ctx.timer_stop()
    if sampler:
        sampler.mark()
    dist.barrier()
    launches = ctx.launches() - l0
    ofs, sz, crc, mo, total = ctx.deflate_results(job, m)
    out = ctx.deflate_fetch(job, total)

    # verification: every stream through zlib, CRCs against zlib.crc32, every stream through the compiled reference
    zs = None
    if method == 93:
        from otezip_b200.zstdlib import Zstd
        zs = Zstd()

    def chk(i):
        p = bytes(out[int(ofs[i]):int(ofs[i]) + int(sz[i])])
        dec = zlib.decompress(p, -15) if mo[i] == 8 else zs.decompress(p, size) if mo[i] == 93 else p
        return dec == srcs[i] and int(crc[i]) == (zlib.crc32(srcs[i]) & 0xFFFFFFFF)
    with cf.ThreadPoolExecutor(NCPU) as ex:
        n_ok = sum(ex.map(chk, range(m)))
    ref_ok, ref_n = None, 0
    try:
        if method == 93:
            raise RuntimeError("the reference rejects real Zstandard frames (SURVEY F3): libzstd is the checker")
        from oracle import RefLib
        ref = RefLib()
        # ZIP32 archives of <= 3 GiB of payload each, read back by the reference on all cores
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        groups, cur, cur_b = [], [], 0
        for i in range(m):
            if cur and cur_b + int(sz[i]) > (3 << 30):
                groups.append(cur)
                cur, cur_b = [], 0
            cur.append(i)
            cur_b += int(sz[i])
        if cur:
            groups.append(cur)
        ref_ok = True
        for g in groups:
            ms_ = [synth.Member("f%d" % i, int(mo[i]), bytes(out[int(ofs[i]):int(ofs[i]) + int(sz[i])]), size, int(crc[i])) for i in g]
            path = os.path.join(tmp, "otz_c5_%d.zip" % os.getpid())
            with open(path, "wb") as f:
                f.write(synth.build_zip(ms_))
            try:
                parts = [list(range(t, len(g), NCPU)) for t in range(NCPU)]

                def rd(ii):
                    err, got = ref.extract_indices(path, ii, verify_crc=1, keep_data=True)
                    return err == 0 and all(got[k] == srcs[g[i]] for k, i in enumerate(ii))
                with cf.ThreadPoolExecutor(NCPU) as ex:
                    ref_ok = ref_ok and all(ex.map(rd, [p for p in parts if p]))
            finally:
                os.unlink(path)
            ref_n += len(g)
    except Exception as e:  # pragma: no cover
        ref_ok = "unavailable: %r" % e
    if n_ok != m or ref_ok is False:
        raise SystemExit("bench c5: %d/%d streams verified by zlib, compiled reference ok=%s" % (n_ok, m, ref_ok))
    # e2e: host buffers through otz_deflate_host
    o_ofs = np.zeros(m, dtype=np.uint64); o_sz = np.zeros(m, dtype=np.uint32); o_crc = np.zeros(m, dtype=np.uint32)
    o_m = np.zeros(m, dtype=np.uint16); tot = C.c_uint64()
    out_host = ctx.pinned(img.nbytes)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def e2e_step():
        ctx.lib.check(ctx.L.otz_deflate_host(ctx.h, vp(img), img.nbytes, vp(in_ofs), vp(in_len), vp(meth), m, vp(out_host), out_host.nbytes,
                                             vp(o_ofs), vp(o_sz), vp(o_crc), vp(o_m), C.byref(tot)), "otz_deflate_host")
    e2e_step()
    dist.barrier(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    ms_max = dist.max(ms)
    tot_in = dist.sum(float(m * size))
    tot_out = dist.sum(float(total))
    e2e_max = dist.max(e2e_s)
    ctx.deflate_destroy(job)
    ctx.dev_free(d_in)
    ctx.pinned_free(img)
    ctx.pinned_free(out_host)
    peak, peak_src = hbm_peak()
    algo = m * size + total
    achieved = algo / (ms / steps / 1e3) / GB
    return {
        "metric": metric_name("c5" if method == 8 else "c5f" if method == 0x108 else "c5z"), "value": tot_in * steps / (ms_max / 1e3) / GB, "unit": "GB/s", "ms_per_step": ms_max / steps,
        "run": {"entries_this_rank": m, "compressed_bytes": int(tot_out), "ratio": tot_in / max(tot_out, 1.0), "reference_ratio_same_level": 4.36,
                "zlib6_ratio": 9.0, "verified": {("zlib_streams_ok" if method != 93 else "libzstd_frames_ok"): n_ok, "compiled_reference_streams_ok": ref_n if ref_ok is True else ref_ok, "of": m}},
        "roofline": {"bound": "hbm", "kernel": "k_deflate_chunks%s (+crc, scan, gather)" % ("" if method == 8 else " -> zse_emit_block"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": int(algo)},
        "e2e": {"value": tot_in * e2e_steps / e2e_max / GB, "unit": "GB/s", "h2d_bytes_per_step": int(img.nbytes),
                "d2h_bytes_per_step": int(total + 18 * m), "steps": e2e_steps,
                "timing": "host wall clock around otz_deflate_host (C-ABI, pinned host buffers), device-synchronised"},
        "gpu_launches": int(launches)}


# ----------------------------------------------------------------------------- configs[0] as written: the CLI
def cli_config0():
    """BASELINE configs[0] is DEFINED as `otezip -x` on a 1,000-entry DEFLATE archive (64 KiB text-like entries) with the
    CRC-32 check (/root/reference/src/main.c:429-585).  Wall clock of the reference CLI (oracle/_ref/otezip_ref, CPU) and of
    the SAME main.c relinked against libotezip_b200.so (oracle/_ref/otezip_relinked), process start, CUDA initialisation and
    the 1,000 file writes included; both trees compared byte for byte."""
    import shutil
    import subprocess
    import tempfile
    from otezip_b200 import synth
    refdir = os.path.join(ROOT, "oracle", "_ref")
    exes = {"reference_cli": os.path.join(refdir, "otezip_ref"), "b200_cli": os.path.join(refdir, "otezip_relinked")}
    if not all(os.path.exists(e) for e in exes.values()):
        return {"error": "oracle/_ref CLIs not built"}
    tmp = tempfile.mkdtemp(prefix="otz_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        z = os.path.join(tmp, "c1.zip")
        ms = synth.config_c1(1000, 65536)
        with open(z, "wb") as f:
            f.write(synth.build_zip(ms))
        total = sum(m.uncomp_size for m in ms)
        res, trees = {}, {}
        for tag, exe in exes.items():
            best = None
            for rep in range(2 if tag == "b200_cli" else 1):     # (the GPU run twice: the first one pages the CUDA libraries in)
                d = os.path.join(tmp, "%s_%d" % (tag, rep))
                os.mkdir(d)
                t0 = time.perf_counter()
                r = subprocess.run([exe, "-x", z, "--verify-crc"], cwd=d, capture_output=True, text=True, timeout=600)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    return {"error": "%s exit %d: %s" % (tag, r.returncode, r.stderr[-300:])}
                best = dt if best is None else min(best, dt)
                tree = {}
                for root, _, fs in os.walk(d):
                    for fn in fs:
                        q = os.path.join(root, fn)
                        tree[os.path.relpath(q, d)] = zlib.crc32(open(q, "rb").read())
                trees[tag] = tree
                shutil.rmtree(d)
            res[tag] = {"seconds": best, "MB_per_s": total / best / 1e6}
        res["identical_trees"] = trees["reference_cli"] == trees["b200_cli"] and len(trees["b200_cli"]) == 1000
        res["speedup"] = res["reference_cli"]["seconds"] / res["b200_cli"]["seconds"]
        res["what"] = "otezip -x c1.zip --verify-crc, 1,000 x 64 KiB (64 MiB) into /dev/shm; wall clock incl. process start, CUDA init and file writes"
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- GPU arm: the read path
def run_extract(ctx, dist: Dist, name: str, n_entries, steps, warmup, e2e_steps, scaling="strong", sampler=None):
    """Device-timed batched extract of one workload + the end-to-end host-buffer call.  -> dict (parts of the JSON line)."""
    import ctypes as C
    wl = workload(name, dist.rank, dist.world, n_entries, ctx.pinned, scaling)
    img, tab = wl["image"], wl["table"]
    n = len(tab)
    d_img = ctx.dev_alloc(img.nbytes)
    ctx.h2d(d_img, img)
    d_out = ctx.dev_alloc(wl["out_bytes"]) if wl["out_bytes"] else None
    plan = ctx.plan(tab, wl["opts"])
    ctx.sync()

    def step():
        ctx.run(plan, d_img, img.nbytes, d_out, wl["out_bytes"])

    for _ in range(warmup):
        step()
    ctx.sync()
    crc, st = ctx.results(plan, n)
    inflate_fallbacks = int(ctx.L.otz_inflate_fallbacks(ctx.h))
    real = (tab["flags"] & 2) == 0          # chunk rows carry no CRC of their own
    ok_mask = 0x200 if name == "c4z" else 0      # real Zstandard frames carry the "reference rejects" flag
    bad = int(np.count_nonzero((st & ~ok_mask) != 0))
    if bad or not np.array_equal(crc[real], tab["crc32"][real]):
        raise SystemExit("bench %s: %d entries failed on the GPU path (status/CRC) — number would be invalid" % (name, bad))

    # ---- device-timed region: K steps, inputs resident in HBM
    dist.barrier()
    ctx.sync()
    ctx.profile(1)
    l0 = ctx.launches()
    if sampler:
        sampler.mark()
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop()
    if sampler:
        sampler.mark()
    dist.barrier()
    launches = ctx.launches() - l0
    prof = ctx.profile_read()
    ctx.profile(0)
    ms_max = dist.max(ms)
    total_uncomp = dist.sum(float(wl["uncomp_bytes"]))
    total_algo = dist.sum(float(wl["algo_bytes"]))
    value = total_uncomp * steps / (ms_max / 1e3) / GB

    # ---- end to end through the C-ABI host-buffer call: pinned host image -> H2D -> kernels -> D2H
    out_host = ctx.pinned(wl["out_bytes"]) if wl["out_bytes"] else None
    L = ctx.L
    crc_h = np.zeros(max(n, 1), dtype=np.uint32)
    st_h = np.zeros(max(n, 1), dtype=np.int32)

    def e2e_step():
        ctx.lib.check(L.otz_extract_host(ctx.h, img.ctypes.data_as(C.c_void_p), img.nbytes, tab.ctypes.data_as(C.c_void_p), n,
                                         C.byref(wl["opts"]), out_host.ctypes.data_as(C.c_void_p) if out_host is not None else None,
                                         wl["out_bytes"], crc_h.ctypes.data_as(C.c_void_p), st_h.ctypes.data_as(C.c_void_p)),
                      "otz_extract_host")
    e2e_step()
    e2e_step()
    dist.barrier()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    dist.barrier()
    assert not np.count_nonzero(st_h[:n] & ~ok_mask) and np.array_equal(crc_h[:n][real], tab["crc32"][real])
    e2e_max = dist.max(e2e_s)
    e2e_value = total_uncomp * e2e_steps / e2e_max / GB
    h2d = dist.sum(float(img.nbytes + tab.nbytes))
    d2h = dist.sum(float(wl["out_bytes"] + 8 * n))

    peak, peak_src = hbm_peak()
    idx = {"resolve": 0, "decode": 1, "crc": 2, "finalize": 3}
    dom = wl["dominant"]
    k_ms = float(np.mean([p[idx[dom]] for p in prof])) if prof else float("nan")
    shares = {k: float(np.mean([p[i] for p in prof])) for k, i in idx.items()} if prof else {}
    achieved = wl["algo_bytes"] / (k_ms / 1e3) / GB
    # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this very
    # configuration (tools/profile_*.sh write the file); null when the entry count differs
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(name)
        if tr and int(tr.get("entries", -1)) == int(wl["n_entries"]) and dist.world == 1:
            traffic, traffic_src = tr["traffic"], "committed ncu --set full capture (profiles/roofline_traffic.json), not measured in this run"
    except Exception:
        pass
    res = {
        "metric": metric_name(name), "value": value, "unit": "GB/s", "ms_per_step": ms_max / steps,
        "run": {"entries_this_rank": int(wl["hi"] - wl["lo"]), "table_rows_this_rank": n, "uncomp_bytes_this_rank": wl["uncomp_bytes"],
                "algorithmic_bytes_this_rank": wl["algo_bytes"], "algorithmic_bytes_all_ranks": total_algo, "inflate_fallbacks": inflate_fallbacks},
        "roofline": {"bound": "hbm", "kernel": {"crc": "k_crc_chunks", "decode": DECODE_KERNELS.get(name, "decode")}[dom],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel_ms": k_ms, "phase_ms": shares,
                     "note": "achieved = algorithmic bytes of this rank (comp + uncomp; configs[1]: uncomp) / event-timed duration of the phase"},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "timing": "host wall clock around otz_extract_host (C-ABI, pinned host buffers; H2D of the archive image and D2H of the "
                          "extracted bytes inside), device-synchronised, max over ranks"},
        "gpu_launches": int(launches),
    }
    ctx.plan_destroy(plan)
    ctx.dev_free(d_img)
    if d_out is not None:
        ctx.dev_free(d_out)
    ctx.pinned_free(img)
    if out_host is not None:
        ctx.pinned_free(out_host)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--entries", type=int, default=None, help="override the entry count (quick runs)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: shard ONE archive (default) or one archive per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configurations (N = 1 default run only)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    dist = Dist(args.gpus)
    if args.impl == "reference":
        run_reference(args, dist)
        dist.close()
        return
    if args.workload == "c3w":
        args.scaling = "weak"   # (written by the GPU compressor of every rank: one archive per rank)
    from otezip_b200 import Ctx
    ctx = Ctx(dist.local)
    # pinned host buffers are placed where the allocating thread runs: keep this rank on the CPUs of its GPU's NUMA node
    numa = None
    try:
        cpus = numa_cpus_of_gpu(ctx.pci_bus_id())
        if cpus:
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = "rank bound to the %d CPUs of its GPU's NUMA node" % len(cpus)
    except Exception:
        pass
    sampler = ClockSampler(ctx.pci_bus_id())
    sampler.start()
    n = args.entries or DEFAULT_ENTRIES[args.workload]
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 5))
    if args.workload in C5_METHOD:
        res = run_c5(ctx, dist, args.entries, args.steps, args.warmup, e2e_steps, sampler, C5_METHOD[args.workload])
    else:
        res = run_extract(ctx, dist, args.workload, args.entries, args.steps, args.warmup, e2e_steps, args.scaling, sampler)
    clocks = sampler.summary()
    line = {
        "metric": res["metric"], "value": res["value"], "unit": "GB/s", "n_gpus": dist.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": wl_config(args.workload, n, dist.world, args.scaling),
        "run": res["run"], "roofline": res["roofline"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": clocks,
    }
    if numa:
        line["run"]["host_placement"] = numa
    if dist.world == 1 and dist.rank == 0:
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args.workload, args.entries)
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % e}
        if not args.no_secondary and args.workload == "c3" and args.entries is None:
            # the other BASELINE configurations, same process, fewer steps: nothing the headline change would hide
            sec = {}
            for w in ("c1", "c2", "c4", "c4z", "c5", "c5f", "c5z"):
                try:
                    r = (run_c5(ctx, dist, None, 3, 3, 1, None, C5_METHOD[w]) if w in C5_METHOD else run_extract(ctx, dist, w, None, 5, 3, 2))
                    sec[w] = {"workload": WORKLOADS[w], "metric": r["metric"], "value": r["value"], "unit": "GB/s", "ms_per_step": r["ms_per_step"],
                              "roofline_frac": r["roofline"]["frac"], "roofline_kernel": r["roofline"]["kernel"], "e2e": r["e2e"]["value"],
                              "run": r["run"]}
                except BaseException as e:  # a failing secondary must not take the headline with it
                    sec[w] = {"workload": WORKLOADS[w], "error": repr(e)}
            try:
                sec["c1_cli"] = cli_config0()
            except BaseException as e:
                sec["c1_cli"] = {"error": repr(e)}
            line["secondary"] = sec
    sampler.stop()
    if dist.rank == 0:
        print(json.dumps(line))
    ctx.close()
    dist.close()


if __name__ == "__main__":
    main()
