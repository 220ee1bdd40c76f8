"""Public-API timing on BASELINE config 1 (the case the reference can run): device-resident and host inputs."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytemdiags_b200 import synthetic as syn, TEMDiagnostics
lat, lon = syn.pg2_grid(30); K, T, L = 72, 24, 50
plev = syn.default_plev(K)
f = syn.synth_fields(lat, lon, plev, T, seed=0)
names = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')
dv = {k: torch.as_tensor(v).cuda() for k, v in f.items()}
for label, src in (('device tensors', dv), ('pageable numpy', f)):
    ts = []
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.time()
        tem = TEMDiagnostics(src['ua'], src['va'], src['ta'], src['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0)
        outs = [getattr(tem, n)() for n in names]
        torch.cuda.synchronize(); ts.append((time.time() - t0) * 1e3)
    print('config1 %-16s ms per call: %s  -> %.3e col*lev*steps/s' % (label, [round(x, 1) for x in ts], lat.shape[0] * K * T / (min(ts) * 1e-3)), flush=True)
