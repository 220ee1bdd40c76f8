"""torchrun --nproc-per-node N tools/sharded_check.py : ShardedTEM over N real GPUs vs the unsharded result.

Checks, bit for bit: `gather_all()` (ONE all-gather for the ten public outputs + the tracer diagnostics) through
torch.distributed/NCCL and through libtemd's own NCCL call (`TemdComm` -> temd_allgather_outputs), with uneven time
slabs (T = 7) and, when N > 7 or with --empty, ranks that own an EMPTY slab."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pytemdiags_b200 import TEMDiagnostics, synthetic as syn
from pytemdiags_b200.distributed import PUBLIC_OUTPUTS, TRACER_PUBLIC, ShardedTEM, TemdComm, h2d_weights, shard_bounds

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = 'cuda:%d' % local
dist.init_process_group('nccl', device_id=torch.device(dev))
lat, lon = syn.pg2_grid(12)
K, L = 9, 40
ok = True
for T in ((7, 1) if '--empty' in sys.argv or world > 7 else (7,)):     # 7: uneven slabs; 1: empty slabs on ranks > 0
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=33, fields=('ua', 'va', 'ta', 'wap', 'q'))
    a, b = shard_bounds(T, world)[rank]
    kw = dict(L=L, dims=('time', 'lev', 'ncol'), debug_level=0, device=dev)
    full = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, q=f['q'], **kw)
    ids = [TemdComm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = TemdComm(world, rank, ids[0], dev)
    wts = [1.0 + (r % 2) for r in range(world)]          # bandwidth-weighted slabs: odd ranks take twice as many steps
    for transport, c, w in (('torch.distributed', None, None), ('libtemd nccl', comm, None), ('torch.distributed weighted', None, wts)):
        a, b = shard_bounds(T, world, w)[rank]
        sh = ShardedTEM(f['ua'][a:b], f['va'][a:b], f['ta'][a:b], f['wap'][a:b], plev, lat, q=f['q'][a:b], T=T, comm=c,
                        weights=w, **kw)
        out = sh.gather_all()
        good = sh.T == T
        for n in PUBLIC_OUTPUTS:
            good &= bool(np.array_equal(out[n].cpu().numpy(), getattr(full, n)()))
        for n in TRACER_PUBLIC:
            good &= bool(np.array_equal(out[n][0].cpu().numpy(), getattr(full, n)(0)))
        good &= bool(np.array_equal(sh.gather('psitem').cpu().numpy(), full.psitem()))
        print('rank %d/%d T=%d slab [%d,%d) %s: sharded == unsharded: %s' % (rank, world, T, a, b, transport, good), flush=True)
        ok &= good
    comm.close()
hw = h2d_weights(dev, nbytes=64 << 20)                  # per-GPU concurrent H2D GB/s, identical list on every rank
ok &= len(hw) == world and all(x > 0 for x in hw)
if rank == 0:
    print('h2d_weights (GB/s per GPU, all ranks copying at once):', [round(x, 1) for x in hw], flush=True)
flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('ALL %d RANKS OK: %s' % (world, bool(flag.item())), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
