"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard -> decode locally -> gather equals the unsharded result."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.bindings import Oracle
    from openairinterface5g_b200.shard import shard_range
    from common import make_case
    orc = Oracle()
    n = 7
    K, P, llr = make_case(orc, 2, 32, 13, n, 3.0, seed=3)       # every rank builds the same seeded workload
    lo, hi = shard_range(n, rank, world)
    mine = np.zeros((n, 32 * 32 // 8 + 1), dtype=np.int64)
    for i in range(lo, hi):
        it, out = orc.decode(2, 32, 13, 8, llr[i])
        mine[i, 0] = it
        mine[i, 1:] = out
    t = torch.from_numpy(mine)
    dist.all_reduce(t)                                           # disjoint shards: sum == gather
    # constant tables broadcast from rank 0 (init-time only)
    blob = torch.from_numpy(np.frombuffer(open(os.path.join(ROOT, "openairinterface5g_b200", "csrc", "nr_bg_tables.h"), "rb").read(), dtype=np.uint8).copy())
    b0 = blob.clone()
    dist.broadcast(b0, src=0)
    ok_tables = bool(torch.equal(blob, b0))
    if rank == 0:
        full = np.zeros_like(mine)
        for i in range(n):
            it, out = orc.decode(2, 32, 13, 8, llr[i])
            full[i, 0] = it
            full[i, 1:] = out
        q.put((bool(np.array_equal(full, t.numpy())), ok_tables))
    dist.destroy_process_group()


def test_shard_decode_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    same, tables = q.get(timeout=5)
    assert same and tables


def test_shard_range_properties():
    from openairinterface5g_b200.shard import shard_range, sticky_gpu
    for n in (0, 1, 7, 1024, 1025):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    assert all(sticky_gpu(u, r, 8) == sticky_gpu(u, r, 8) and 0 <= sticky_gpu(u, r, 8) < 8 for u in range(16) for r in range(34))
