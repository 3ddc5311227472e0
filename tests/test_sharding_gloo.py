"""The N>1 path of bench.py on CPU: world_size-2 gloo rendezvous, stream sharding with no
data-path collective, barrier + max-over-ranks time and whole-job sums."""
import os
import socket
import sys

import pytest

from common import ROOT


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import bench
    dd = bench.Dist(world)
    mine = bench.shard_streams(8, dd.rank, dd.world)
    dd.barrier()
    slowest = dd.max(1.0 + rank)          # max over ranks, as for ms_per_step
    total = dd.sum(float(len(mine)))      # whole-job units
    q.put((rank, mine, slowest, total))
    dd.close()


def test_world_size_two_gloo():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    assert all(r[2] == 2.0 and r[3] == 8.0 for r in res)


def test_shard_covers_all_streams_once():
    import bench
    for world in (1, 2, 4, 8):
        seen = sorted(s for r in range(world) for s in bench.shard_streams(64, r, world))
        assert seen == list(range(64))
        assert all(len(bench.shard_streams(64, r, world)) == 64 // world for r in range(world))
