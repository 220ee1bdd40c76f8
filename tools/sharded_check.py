"""torchrun --nproc-per-node N tools/sharded_check.py : ShardedTEM over N real GPUs (NCCL) vs the unsharded result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pytemdiags_b200 import TEMDiagnostics, synthetic as syn
from pytemdiags_b200.distributed import ShardedTEM, shard_bounds
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
lat, lon = syn.pg2_grid(12); K, T, L = 9, 7, 40          # T = 7 is not divisible by 2/4/8: uneven slabs
plev = syn.default_plev(K)
f = syn.synth_fields(lat, lon, plev, T, seed=33)
a, b = shard_bounds(T, world)[rank]
sh = ShardedTEM(f['ua'][a:b], f['va'][a:b], f['ta'][a:b], f['wap'][a:b], plev, lat, T=T, L=L, dims=('time', 'lev', 'ncol'),
                debug_level=0, device='cuda:%d' % local) if b > a else None
full = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0,
                      device='cuda:%d' % local)
ok = True
for n in ('vtem', 'epfy', 'epdiv', 'psitem', 'utendwtem'):
    got = sh.gather(n).cpu().numpy()
    ok &= bool(np.array_equal(got, getattr(full, n)()))
print('rank %d/%d slab [%d,%d) sharded == unsharded: %s' % (rank, world, a, b, ok), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
