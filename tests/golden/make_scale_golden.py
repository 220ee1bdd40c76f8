"""Expected outputs for the full-size BASELINE configs, computed with the CPU oracle in the
build container so that the GPU box only has to run the CUDA path and compare.

Inputs are NOT stored: `pytemdiags_b200.synthetic.synth_fields` regenerates them bit-identically from
(grid, K, seed, t0).  The oracle here is the factored form with the recurrence basis and the normal-equation
inverse (the literal N x N operator needs 1-20 TB at these grids); both are pinned against SciPy / the
unmodified reference at small sizes by tests/test_oracle_golden.py.

    python tests/golden/make_scale_golden.py      # ~10 minutes on 8 cores, writes tests/golden/scale_*.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from pytemdiags_b200 import synthetic as syn  # noqa: E402

KEEP = oracle.TEM_OUTPUTS + ('ub', 'vb', 'thetab', 'wapb', 'upvpb', 'upwappb', 'vptpb')
TEM_CASES = {
    'scale_config2_t1': dict(grid=('pg2', 120), K=72, L=100, seed=1, t0=182),
    'scale_config3_t1': dict(grid=('pg2', 256), K=128, L=200, seed=2, t0=48),
    'scale_config4_t1': dict(grid=('latlon', 721, 1440), K=37, L=300, seed=3, t0=120),
}
SWEEP_L = (25, 50, 100, 200, 400, 800)


def main():
    for name, c in ([] if '--sweep-only' in sys.argv else TEM_CASES.items()):
        t0 = time.time()
        lat, lon = syn.make_grid(c['grid'])
        plev = syn.default_plev(c['K'])
        f = syn.synth_fields(lat, lon, plev, 1, seed=c['seed'], t0=c['t0'])
        mats = oracle.sph_matrices(lat, oracle.zm_latitudes(1), c['L'], method='normal', basis='recurrence')
        tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0))
        ref = oracle.tem_suite(tr(f['ua']), tr(f['va']), tr(f['ta']), tr(f['wap']), plev, lat, L=c['L'], literal=False, matrices=mats)
        out = {k: ref[k] for k in KEEP}
        out.update(grid=np.array([str(x) for x in c['grid']]), K=c['K'], L=c['L'], seed=c['seed'], t0=c['t0'])
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'done in %.0f s' % (time.time() - t0), flush=True)
    # config 5: zonal-mean-only L sweep on ne120pg2 x 72 levels, one time step
    lat, lon = syn.pg2_grid(120)
    lat_out = oracle.zm_latitudes(1)
    f = syn.synth_fields(lat, lon, syn.default_plev(72), 1, seed=4, fields=('ua',))['ua'][0]     # [K][N]
    out = {}
    for L in SWEEP_L:
        t0 = time.time()
        Y0, Y0inv, Y0p = oracle.sph_matrices(lat, lat_out, L, method='normal', basis='recurrence')
        c = Y0inv @ f.T
        out['zm_L%d' % L] = (Y0p @ c).T                      # [K][M]
        out['zn_L%d' % L] = (Y0[:512] @ c).T                 # native mean at the first 512 columns
        print('sweep L', L, '%.0f s' % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, 'scale_config5_sweep.npz'), **out)


if __name__ == '__main__':
    main()
