"""cProfile of ONE public-API call on pinned host arrays (config-2 grid, 16 steps): where does the host time go?"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pytemdiags_b200 import TEMDiagnostics, synthetic as syn

lat, lon = syn.pg2_grid(120)
K, T, L = 72, 16, 100
plev = syn.default_plev(K)
N = lat.shape[0]
host = []
g = torch.Generator().manual_seed(0)
for fi in range(4):
    h = torch.empty((T, K, N), dtype=torch.float64).pin_memory()
    h.normal_(generator=g)
    if fi == 2:
        h.mul_(0.5).add_(250.0 + 60.0 * torch.linspace(0, 1, K).view(1, K, 1))
    host.append(h.numpy())
names = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')


def call():
    tem = TEMDiagnostics(host[0], host[1], host[2], host[3], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0)
    return [getattr(tem, n)() for n in names]


for _ in range(3):
    call()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    call()
torch.cuda.synchronize()
print('ms per call %.1f (pure H2D at 55.6 GB/s: %.1f)' % ((time.time() - t0) / 5 * 1e3, 4 * T * K * N * 8 / 55.6e9 * 1e3))
pr = cProfile.Profile()
pr.enable()
call()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
