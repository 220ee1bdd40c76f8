import sys, time, os, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytemdiags_b200 import synthetic as syn, TEMDiagnostics
from pytemdiags_b200.engine import Engine
cfg = syn.CONFIGS['config2']; K = cfg['K']; L = 100; Te = 16
lat, lon = syn.make_grid(cfg['grid']); N = lat.shape[0]; plev = syn.default_plev(K)
dev = torch.device('cuda:0')
eng = Engine(lat, np.arange(-89.5, 90, 1.0), L, device=dev).build_basis()
latr, lonr, plev_d = eng._dev(np.deg2rad(lat)), eng._dev(np.deg2rad(lon)), eng._dev(plev)
big = [eng.synth_fields(fi, 0, 0, 73, plev, latr, lonr, plev_d) for fi in range(4)]   # 58 GB resident like bench.py
host = []
for fi in range(4):
    h = torch.empty((Te, K, N), dtype=torch.float64).pin_memory(); h.copy_(big[fi][:Te * K].reshape(Te, K, N)); host.append(h.numpy())
torch.cuda.synchronize()
names = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')
for mode in ('default', 'gc_disabled'):
    if mode == 'gc_disabled': gc.disable()
    ts = []
    for rep in range(14):
        torch.cuda.synchronize(); t0 = time.time()
        tem = TEMDiagnostics(host[0], host[1], host[2], host[3], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0, device=dev)
        outs = [getattr(tem, n)() for n in names]
        torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
    print(mode, [round(x) for x in ts], flush=True)
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
