"""CPU, world_size 2, gloo: the time-slab sharding and the final all-gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from pytemdiags_b200.distributed import shard_bounds
    assert shard_bounds(96, 8) == [(12 * r, 12 * r + 12) for r in range(8)]
    b = shard_bounds(365, 8)
    assert b[0] == (0, 46) and b[-1][1] == 365 and sum(y - x for x, y in b) == 365
    assert max(y - x for x, y in b) - min(y - x for x, y in b) == 1
    assert all(b[i][1] == b[i + 1][0] for i in range(7))
    assert shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]


def _worker(rank, world, port, T, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pytemdiags_b200.distributed import gather_time_sharded, shard_bounds
    full = torch.arange(5 * 3 * T, dtype=torch.float64).reshape(5, 3, T)     # (lat, plev, time)
    a, b = shard_bounds(T, world)[rank]
    got = gather_time_sharded(full[:, :, a:b].contiguous(), T)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize('T', [8, 7, 1])
def test_gather_time_sharded_gloo(T):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
