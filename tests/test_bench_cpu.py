"""Host-side logic of bench.py that needs no GPU: archive-set builder, distributed plumbing (gloo, world 2),
reference arm."""
import json
import os
import socket
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_archive_set_builder_matches_central_directory_walk(oracle):
    import bench
    from otezip_b200.native import parse_central

    def gen(i):
        d = np.random.default_rng(i).bytes(1000 + 37 * i)
        return d, len(d), zlib.crc32(d) & 0xFFFFFFFF
    img, tab, out_bytes = bench.build_archive_set(lambda n: np.zeros(n, dtype=np.uint8), gen, range(25), 10, 0, threads=2)
    # three ZIP32 files back to back; the first one, cut out, is a valid archive on its own
    first_len = int(tab["lfh_ofs"][10])
    one = bytes(img[:first_len])
    eocd = one.rfind(b"PK\x05\x06")
    t1 = parse_central(one[:eocd + 22])
    assert len(t1) == 10
    assert np.array_equal(t1["crc32"], tab["crc32"][:10]) and np.array_equal(t1["lfh_ofs"], tab["lfh_ofs"][:10])
    rc, ents = oracle.load_central(one[:eocd + 22])
    assert rc == 0 and [e.crc32 for e in ents] == [int(c) for c in tab["crc32"][:10]]
    assert out_bytes == int(((tab["uncomp_size"].astype(np.int64) + 15) & ~15).sum())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_reference_arm_under_torchrun_world2():
    """N>1 path of the reference arm: rank 0 alone prints the JSON line, the other rank exits 0 (gloo on CPU)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--workload", "c4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "GB/s"
    # the reference arm names the workload exactly as the GPU arm would (the driver compares the two `config`s)
    import bench
    assert d["config"] == bench.wl_config("c4", 10000, 2, "strong") and d["scaling"] == "strong"


def test_partition_balances_bytes_in_contiguous_ranges():
    """otz_partition: contiguous index ranges, balanced by comp + uncomp bytes — host code, no GPU."""
    from otezip_b200 import synth
    from otezip_b200.native import partition, ENTRY_DTYPE
    sizes = synth.config_c3_sizes(10000, 3)
    t = np.zeros(len(sizes), dtype=ENTRY_DTYPE)
    t["uncomp_size"] = sizes
    t["comp_size"] = np.array(sizes) // 9
    w = t["uncomp_size"].astype(np.int64) + t["comp_size"]
    for parts in (1, 2, 3, 4, 8):
        f = partition(t, parts)
        assert f[0] == 0 and f[-1] == len(t) and np.all(np.diff(f.astype(np.int64)) >= 0)
        loads = [int(w[f[g]:f[g + 1]].sum()) for g in range(parts)]
        assert max(loads) <= sum(loads) / parts + int(w.max())      # never worse than one entry off
    # degenerate tables: fewer entries than parts, empty table, one giant entry
    f = partition(t[:3], 8)
    assert f[0] == 0 and f[-1] == 3 and np.all(np.diff(f.astype(np.int64)) >= 0)
    assert list(partition(t[:0], 4)) == [0, 0, 0, 0, 0]
    g = t[:5].copy()
    g["uncomp_size"] = [10, 10, 1 << 30, 10, 10]
    f = partition(g, 2)
    assert f[0] == 0 and f[2] == 5 and 2 <= f[1] <= 3


def test_shards_of_one_archive_cover_it_exactly():
    """bench.py --gpus N: the ranks' shards of configs[2] are disjoint, contiguous and cover the entry list."""
    import bench
    pl = bench.entry_plan("c3", 10000)
    for world in (2, 4, 8):
        edges = [bench.shard_of("c3", pl["sizes"], world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 10000
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        loads = [sum(pl["sizes"][a:b]) for a, b in edges]
        assert max(loads) / (sum(loads) / world) < 1.01


def _shard_worker(rank, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import bench
    d = bench.Dist(2)
    # a small configs[2]-shaped archive, sharded: every rank builds only its own byte range
    wl = bench.workload("c3", d.rank, d.world, 40, lambda n: np.zeros(n, dtype=np.uint8))
    tab = wl["table"]
    total = d.sum(float(wl["uncomp_bytes"]))
    n_all = d.sum(float(len(tab)))
    ok = all(zlib.crc32(zlib.decompress(bytes(wl["image"][int(t["lfh_ofs"]) + 30 + 11:int(t["lfh_ofs"]) + 30 + 11 + int(t["comp_size"])]), -15))
             == int(t["crc32"]) for t in tab)
    q.put((rank, wl["lo"], wl["hi"], total, n_all, ok))
    d.close()


def test_sharded_workload_gloo_world2():
    import multiprocessing as mp
    import bench
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_shard_worker, args=(r, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    pl = bench.entry_plan("c3", 40)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 40
    assert res[0][3] == res[1][3] == float(sum(pl["sizes"])) and res[0][4] == 40.0
    assert res[0][5] and res[1][5]


def _dist_worker(rank, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import bench
    d = bench.Dist(2)
    d.barrier()
    q.put((rank, d.max(float(rank + 1)), d.sum(10.0 * (rank + 1))))
    d.close()


def test_dist_max_and_sum_gloo_world2():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_dist_worker, args=(r, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, 2.0, 30.0), (1, 2.0, 30.0)]
