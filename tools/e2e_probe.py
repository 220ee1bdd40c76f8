"""Times the phases of one public-API call on pinned host inputs (diagnostic for the e2e number)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytemdiags_b200 import synthetic as syn, TEMDiagnostics
from pytemdiags_b200.engine import Engine
cfg = syn.CONFIGS['config2']; K = cfg['K']; L = 100; Te = 16
lat, lon = syn.make_grid(cfg['grid']); N = lat.shape[0]; plev = syn.default_plev(K)
dev = torch.device('cuda:0')
eng = Engine(lat, np.arange(-89.5, 90, 1.0), L, device=dev).build_basis()
latr, lonr, plev_d = eng._dev(np.deg2rad(lat)), eng._dev(np.deg2rad(lon)), eng._dev(plev)
host = []
for fi in range(4):
    x = eng.synth_fields(fi, 0, 0, Te, plev, latr, lonr, plev_d)
    h = torch.empty((Te, K, N), dtype=torch.float64).pin_memory(); h.copy_(x.reshape(Te, K, N)); host.append(h.numpy())
torch.cuda.synchronize()
def sync(): torch.cuda.synchronize(); return time.time()
for slab in (1 << 30, 2 << 30, 4 << 30, 16 << 30):
    for rep in range(3):
        t0 = sync()
        tem = TEMDiagnostics(host[0], host[1], host[2], host[3], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0, device=dev, slab_bytes=slab)
        t1 = sync()
        outs = [getattr(tem, n)() for n in ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')]
        t2 = sync()
    print('slab %5.1f GB: ctor %.1f ms, outputs %.1f ms' % (slab / 2**30, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
# raw copies, same chunking as the 2 GB slab path
t0 = sync()
for t in range(0, Te, 2):
    for h in host:
        torch.from_numpy(h[t:t + 2]).to(dev, non_blocking=True)
t1 = sync()
print('raw chunked H2D: %.1f ms (%.1f GB/s)' % ((t1 - t0) * 1e3, 4 * Te * K * N * 8 / (t1 - t0) / 1e9))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
tem = TEMDiagnostics(host[0], host[1], host[2], host[3], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0, device=dev)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
