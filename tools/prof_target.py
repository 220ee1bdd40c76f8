"""One launch of each hot kernel on a BASELINE slab, for `ncu -k regex:<kernel> --launch-skip 1 -c 1 --set full`.
  python tools/prof_target.py <config> <slab_steps> [dedup]      (runs every kernel twice: warm-up + target)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pytemdiags_b200 import constants as const, synthetic as syn
from pytemdiags_b200.engine import DedupEngine, Engine

name, Ts = sys.argv[1], int(sys.argv[2])
dedup = len(sys.argv) > 3 and sys.argv[3] == 'dedup'
cfg = syn.CONFIGS[name]
L, K = cfg['L'], cfg['K']
lat, lon = syn.make_grid(cfg['grid'])
plev = syn.default_plev(K)
lat_zm = np.arange(-89.5, 90, 1.0)
eng = (DedupEngine if dedup else Engine)(lat, lat_zm, L, device=torch.device('cuda:0'))
eng.build_basis()
latr, lonr, plev_d = eng._dev(np.deg2rad(lat)), eng._dev(np.deg2rad(lon)), eng._dev(plev)
xs = [eng.synth_fields(fi, 0, 0, Ts, plev, latr, lonr, plev_d) for fi in range(4)]
lev_scale = eng._dev((const.P0 / (plev * 100)) ** const.k)
f_zm = 2 * const.Om * np.sin(lat_zm * np.pi / 180)
coslat = np.cos(lat_zm * np.pi / 180)
for _ in range(2):
    c4, cf = eng.tem_coefficients(xs, lev_scale, K)
    coef = torch.cat([c4, cf], 0)
    zm = eng.synth_out(coef).reshape(7, Ts, K, eng.M)
    res = eng.tem_epilogue(zm, plev * 100, f_zm, coslat)
torch.cuda.synchronize()
print('done', name, Ts, 'dedup' if dedup else 'dense')
