"""e2e with ordinary (pageable) numpy inputs vs pinned ones."""
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
pinned, pageable = [], []
for fi in range(4):
    x = eng.synth_fields(fi, 0, 0, Te, plev, latr, lonr, plev_d)
    h = torch.empty((Te, K, N), dtype=torch.float64).pin_memory(); h.copy_(x.reshape(Te, K, N)); pinned.append(h.numpy())
    pageable.append(h.numpy().copy())
torch.cuda.synchronize()
for name, host in (('pinned', pinned), ('pageable', pageable)):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.time()
        tem = TEMDiagnostics(host[0], host[1], host[2], host[3], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0, device=dev)
        v = tem.vtem()
        torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
    print(name, [round(x) for x in ts], 'GB/s %.1f' % (4 * Te * K * N * 8 / (min(ts) * 1e-3) / 1e9), flush=True)
# reference-internal layout (ncol, plev, time), pageable
ref_layout = [np.ascontiguousarray(h.transpose(2, 1, 0)) for h in pageable]
ts = []
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    tem = TEMDiagnostics(ref_layout[0], ref_layout[1], ref_layout[2], ref_layout[3], lat, p=plev, L=L, debug_level=0, device=dev)
    v = tem.vtem()
    torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
print('(ncol, plev, time) pageable', [round(x) for x in ts], flush=True)
