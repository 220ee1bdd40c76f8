"""Kernel A/B timings on one GPU: k_eddy / k_project on the BASELINE slabs under the tuning knobs
(TEMD_EDDY_WARPS, TEMD_EDDY_NCH are read at every launch).  python tools/kbench.py [config2 config3 config4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pytemdiags_b200 import constants as const, synthetic as syn
from pytemdiags_b200.engine import Engine

PEAK = 37.1
SLABS = {'config2': 73, 'config3': 12, 'config4': 30, 'config1': 24}
dev = torch.device('cuda:0')


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name in (sys.argv[1:] or ['config2', 'config3', 'config4']):
    cfg = syn.CONFIGS[name]
    L, K, Ts = cfg['L'], cfg['K'], SLABS[name]
    lat, lon = syn.make_grid(cfg['grid'])
    N = lat.shape[0]
    plev = syn.default_plev(K)
    eng = Engine(lat, np.arange(-89.5, 90, 1.0), L, device=dev)
    eng.build_basis()
    latr, lonr, plev_d = eng._dev(np.deg2rad(lat)), eng._dev(np.deg2rad(lon)), eng._dev(plev)
    xs = [eng.synth_fields(fi, 0, 0, Ts, plev, latr, lonr, plev_d) for fi in range(4)]
    lev_scale = eng._dev((const.P0 / (plev * 100)) ** const.k)
    rows, Lp = Ts * K, L + 1
    c4 = eng.project(xs, lev_scale=lev_scale, scale_field=2, nlev=K)
    tp = timeit(lambda: eng.project(xs, lev_scale=lev_scale, scale_field=2, nlev=K))
    print('%s rows %d lpad %d: project %.2f ms %.2f TF (%.3f)' % (name, rows, eng.lpad, tp, 8.0 * Lp * N * rows / tp / 1e9,
                                                                 8.0 * Lp * N * rows / tp / 1e9 / PEAK), flush=True)
    ref = None
    for mode, warps, nch, gb in (('fused', 16, 2, 16), ('fused', 8, 2, 16), ('fused', 12, 2, 16), ('split', 0, 0, 16)):
        os.environ['TEMD_EDDY_SCRATCH_GB'] = str(gb)
        os.environ['TEMD_EDDY_BM'] = '24' if warps == 12 else '32'      # 12 warps = the BM = 24 layout (L + 1 <= 104 only)
        os.environ['TEMD_EDDY_MODE'], os.environ['TEMD_EDDY_WARPS'], os.environ['TEMD_EDDY_NCH'] = mode, str(warps), str(nch)
        try:
            cf = eng.eddy_flux_project(xs[0], xs[1], xs[2], xs[3], c4, lev_scale, K)
            te = timeit(lambda: eng.eddy_flux_project(xs[0], xs[1], xs[2], xs[3], c4, lev_scale, K))
        except Exception as e:      # noqa: BLE001
            print('  warps %d nch %d: %s' % (warps, nch, str(e)[:100]))
            continue
        same = '' if ref is None else ' maxdiff vs first %.2e' % float((cf - ref).abs().max() / ref.abs().max())
        ref = cf if ref is None else ref
        tf = 14.0 * Lp * N * rows / te / 1e9
        print('  eddy_flux_project %s warps %2d nch %d scratch %d GB: %.2f ms %.2f TF (%.3f)%s' % (mode, warps, nch, gb, te, tf, tf / PEAK, same), flush=True)
    for v in ('TEMD_EDDY_MODE', 'TEMD_EDDY_WARPS', 'TEMD_EDDY_NCH', 'TEMD_EDDY_SCRATCH_GB', 'TEMD_EDDY_BM'):
        os.environ.pop(v)
    nat = torch.empty((rows, N + (N & 1)), dtype=torch.float64, device=dev)
    tn = timeit(lambda: eng.synth_native(c4[0], out=nat))
    print('  synth_native (1 field): %.2f ms %.2f TF (%.3f)' % (tn, 2.0 * Lp * N * rows / tn / 1e9, 2.0 * Lp * N * rows / tn / 1e9 / PEAK), flush=True)
    del nat
    del eng, xs, c4
    torch.cuda.empty_cache()
