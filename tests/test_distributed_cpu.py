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
    # weighted slabs (heterogeneous host links): contiguous, cover [0, T), proportional to the weights
    w = [23.3] * 4 + [35.4] * 4
    b = shard_bounds(128, 8, w)
    assert b[0][0] == 0 and b[-1][1] == 128 and all(b[i][1] == b[i + 1][0] for i in range(7))
    n = [y - x for x, y in b]
    assert all(abs(n[i] - 128 * w[i] / sum(w)) <= 1 for i in range(8))
    assert n == [13, 13, 13, 13, 19, 19, 19, 19]            # minimax: max steps/weight = 13/23.3, not 20/35.4
    assert shard_bounds(10, 2, [1.0, 1.0]) == [(0, 5), (5, 10)]
    assert shard_bounds(1, 2, [1.0, 2.0]) == [(0, 0), (0, 1)]
    with pytest.raises(ValueError):
        shard_bounds(10, 2, [1.0, -1.0])


def _worker(rank, world, port, Ts, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ok = all(_gather_case(rank, world, T) for T in Ts) and all(_sharded_case(rank, world, T) for T in Ts)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _gather_case(rank, world, T):
    from pytemdiags_b200.distributed import gather_time_major, gather_time_sharded, shard_bounds
    full = torch.arange(5 * 3 * T, dtype=torch.float64).reshape(5, 3, T)     # (lat, plev, time)
    a, b = shard_bounds(T, world)[rank]
    got = gather_time_sharded(full[:, :, a:b].contiguous(), T)
    ok = bool(torch.equal(got, full))
    # the product path: ONE collective for a stack of planes [P][T_local][lev][lat], empty slabs included
    planes = torch.arange(4 * T * 3 * 5, dtype=torch.float64).reshape(4, T, 3, 5)
    got2 = gather_time_major(planes[:, a:b].contiguous(), T)
    ok = ok and bool(torch.equal(got2, planes)) and tuple(got2.shape) == (4, T, 3, 5)
    w = [1.0, 3.0]                                              # weighted slabs: rank 1 owns three quarters
    a, b = shard_bounds(T, world, w)[rank]
    got3 = gather_time_major(planes[:, a:b].contiguous(), T, weights=w)
    return ok and bool(torch.equal(got3, planes))


def _run(target, world, *args):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return sorted(res)


def test_gather_and_sharded_tem_gloo():
    """world_size 2, gloo: even slabs, T % world != 0, and an EMPTY slab on rank 1 (T = 1), for the raw gathers and
    for ShardedTEM's host logic (one process pair runs every case: spawning + importing torch dominates)."""
    assert _run(_worker, 2, (8, 7, 1)) == [(0, True), (1, True)]


def _sharded_case(rank, world, T):
    """ShardedTEM host logic on CPU: the local TEMDiagnostics is replaced by a stub that holds the device-layout
    result planes, so the slab bookkeeping, the empty-slab path and the single collective run under gloo."""
    import types
    import pytemdiags_b200.tem as tem_mod
    from pytemdiags_b200 import distributed as D
    K, M = 3, 6
    names = D.PUBLIC_OUTPUTS
    truth = {n: torch.arange(T * K * M, dtype=torch.float64).reshape(T, K, M) + 1000.0 * i for i, n in enumerate(names)}
    a, b = D.shard_bounds(T, world)[rank]

    class Stub:
        def __init__(self, ua, *args, **kw):
            self.NT, self.NLEV, self.ZM_N, self.ntrac = ua.shape[0], K, M, 0
            self.ZM = types.SimpleNamespace(_engine=types.SimpleNamespace(device=torch.device('cpu')))
            self._dev_results = {n: truth[n][a:b] for n in names}
            self._dev_tracer = []
    orig = tem_mod.TEMDiagnostics
    tem_mod.TEMDiagnostics = Stub
    try:
        ua = np.zeros((b - a, K, 10))
        sh = D.ShardedTEM(ua, ua, ua, ua, None, None, T=None, dims=('time', 'lev', 'ncol'), device='cpu')
        ok = sh.T == T and (sh.local is None) == (b == a) and (sh.K, sh.M) == (K, M)
        out = sh.gather_all()
        for n in names:
            ok = ok and bool(torch.equal(out[n], truth[n].permute(2, 1, 0)))
        ok = ok and bool(torch.equal(sh.gather('epfy'), truth['epfy'].permute(2, 1, 0)))
        dev_layout = sh.gather_all(('vtem',), layout='device')['vtem']
        ok = ok and bool(torch.equal(dev_layout, truth['vtem']))
    finally:
        tem_mod.TEMDiagnostics = orig
    return bool(ok)
