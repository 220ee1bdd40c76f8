"""GPU: the time-sharded driver on one rank (NCCL, world_size 1); the 2-rank logic is covered on CPU with gloo."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sharded_tem_single_rank_nccl():
    import torch
    import torch.distributed as dist
    from pytemdiags_b200 import TEMDiagnostics, synthetic as syn
    from pytemdiags_b200.distributed import ShardedTEM, shard_bounds
    import socket
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    created = False
    if not dist.is_initialized():
        dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda:0'))
        created = True
    try:
        lat, lon = syn.pg2_grid(6)
        K, T, L = 6, 5, 20
        plev = syn.default_plev(K)
        f = syn.synth_fields(lat, lon, plev, T, seed=21)
        a, b = shard_bounds(T, 1)[0]
        sh = ShardedTEM(f['ua'][a:b], f['va'][a:b], f['ta'][a:b], f['wap'][a:b], plev, lat, T=T, L=L,
                        dims=('time', 'lev', 'ncol'), debug_level=0)
        full = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0)
        for n in ('vtem', 'epdiv', 'psitem'):
            got = sh.gather(n).cpu().numpy()
            assert got.shape == (180, K, T)
            assert np.array_equal(got, getattr(full, n)())
    finally:
        if created:
            dist.destroy_process_group()
