"""GPU: every tile-shape variant of the fused kernels on small problems (compute-sanitizer is closed on the
pool, so coverage of ragged tiles / empty split ranges / tiny and large L comes from parity alone).

k_eddy picks BM = 32 / 16 / 8 rows for L+1 <= 104 / 208 / 408 and 16 or 8 consumer warps; k_project picks
1..13 n8-tiles per l-block and 1..4 l-blocks; rows are deliberately not multiples of the row tiles."""
import numpy as np
import pytest

import oracle
from pytemdiags_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-10
ZM7 = ('ub', 'vb', 'thetab', 'wapb', 'upvpb', 'upwappb', 'vptpb')


def nerr(x, ref):
    return float(np.abs(np.asarray(x) - ref).max() / max(np.abs(ref).max(), 1e-300))


CASES = [
    # ne, K, T, L            what it exercises
    (3, 3, 1, 2),            # L+1 = 3: one n8-tile, most GEMM2 warps idle
    (3, 2, 2, 7),            # L+1 = 8 exactly
    (4, 7, 3, 8),            # L+1 = 9 -> 2 tiles, rows = 21
    (5, 5, 3, 17),           # 3 tiles
    (6, 9, 5, 33),           # 5 tiles, rows = 45 (two 32-row tiles, second ragged)
    (8, 11, 3, 62),          # 8 tiles
    (10, 13, 3, 90),         # 12 tiles
    (12, 6, 5, 103),         # 13 tiles (largest BM = 32 case), rows = 30
    (20, 5, 3, 104),         # 14 tiles -> BM = 16
    (20, 7, 3, 130),         # 17 tiles, rows = 21
    (32, 3, 3, 207),         # 26 tiles (largest BM = 16 case)
    (32, 5, 1, 210),         # 27 tiles -> BM = 8, 8 warps
    (32, 3, 2, 250),         # 32 tiles
]


@pytest.mark.parametrize('ne,K,T,L', CASES)
def test_fused_path_variants(ne, K, T, L):
    import torch
    from pytemdiags_b200.engine import Engine
    lat, lon = syn.pg2_grid(ne)
    N = lat.shape[0]
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=100 + L)
    lat_out = oracle.zm_latitudes(1)
    mats = oracle.sph_matrices(lat, lat_out, L, method='normal' if N > 6000 else 'auto',
                               basis='recurrence' if N > 6000 else 'scipy')
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0))
    ref = oracle.tem_suite(tr(f['ua']), tr(f['va']), tr(f['ta']), tr(f['wap']), plev, lat, L=L, literal=False, matrices=mats)
    eng = Engine(lat, lat_out, L).build_basis()
    C = oracle.CONSTANTS
    sc = torch.as_tensor((C['P0'] / (plev * 100)) ** C['k']).cuda()
    xs = [torch.as_tensor(f[n].reshape(T * K, N)).cuda() for n in ('ua', 'va', 'ta', 'wap')]
    c4 = eng.project(xs, lev_scale=sc, scale_field=2, nlev=K)
    cf = eng.eddy_flux_project(xs[0], xs[1], xs[2], xs[3], c4, sc, K)
    zm = eng.synth_out(torch.cat([c4, cf], 0)).cpu().numpy().reshape(7, T, K, -1)
    for i, n in enumerate(ZM7):
        e = nerr(zm[i].transpose(2, 1, 0), ref[n])
        assert e < TOL, (n, e)
    nat = eng.synth_native(c4[0]).cpu().numpy()
    Y0, Y0inv, _ = mats
    assert nerr(nat, (Y0 @ (Y0inv @ f['ua'].reshape(T * K, N).T)).T) < TOL


@pytest.mark.parametrize('warps', ['8', '16'])
def test_eddy_warp_layouts(monkeypatch, warps):
    """both consumer-warp layouts of k_eddy (the default is 16 for BM >= 16)"""
    monkeypatch.setenv('TEMD_EDDY_WARPS', warps)
    monkeypatch.setenv('TEMD_EDDY_MODE', 'fused')
    test_fused_path_variants(8, 11, 3, 62)
    test_fused_path_variants(20, 7, 3, 130)


def test_eddy_bm24_layout(monkeypatch):
    """The opt-in BM = 24 / 12-warp / 5-stage layout of k_eddy (TEMD_EDDY_BM=24; measured 80.4 vs 78.5 ms for the
    default 8-warp BM = 32 layout on the config-2 slab, so it stays an A/B knob)."""
    monkeypatch.setenv('TEMD_EDDY_BM', '24')
    monkeypatch.setenv('TEMD_EDDY_MODE', 'fused')
    for case in ((3, 3, 1, 2), (4, 7, 3, 8), (6, 9, 5, 33), (10, 13, 3, 90), (12, 6, 5, 103)):
        test_fused_path_variants(*case)


@pytest.mark.parametrize('mode', ['fused', 'split'])
def test_eddy_implementations(monkeypatch, mode):
    """temd_eddy_flux_project has two implementations (fused k_eddy: default for L + 1 <= 104; split k_synth eddy
    epilogue + k_project product mode: default above).  Force each one across the L range both can serve, including
    the fused kernel's BM = 16 / 8 variants that the default no longer selects, ragged row batches (scratch budget of
    one 128-row batch) and the product-mode projection with several l-blocks."""
    monkeypatch.setenv('TEMD_EDDY_MODE', mode)
    monkeypatch.setenv('TEMD_EDDY_SCRATCH_GB', '1e-6')          # split: 128-row batches -> several batches, last one ragged
    for case in ((4, 7, 3, 8), (6, 9, 5, 33), (12, 6, 5, 103), (20, 5, 3, 104), (16, 50, 3, 130), (32, 3, 3, 207),
                 (32, 5, 1, 210), (32, 3, 2, 250)):
        test_fused_path_variants(*case)


def test_split_and_fused_agree_closely():
    """Same contraction, different kernels: the flux coefficients agree to ~1e-13 of their magnitude."""
    import os
    import torch
    from pytemdiags_b200.engine import Engine
    lat, lon = syn.pg2_grid(16)
    K, T, L = 40, 4, 150
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=77)
    eng = Engine(lat, oracle.zm_latitudes(1), L).build_basis()
    C = oracle.CONSTANTS
    sc = torch.as_tensor((C['P0'] / (plev * 100)) ** C['k']).cuda()
    xs = [torch.as_tensor(f[n].reshape(T * K, -1)).cuda() for n in ('ua', 'va', 'ta', 'wap')]
    c4 = eng.project(xs, lev_scale=sc, scale_field=2, nlev=K)
    out = {}
    old = os.environ.get('TEMD_EDDY_MODE')
    try:
        for mode in ('fused', 'split'):
            os.environ['TEMD_EDDY_MODE'] = mode
            out[mode] = eng.eddy_flux_project(xs[0], xs[1], xs[2], xs[3], c4, sc, K)
    finally:
        if old is None:
            os.environ.pop('TEMD_EDDY_MODE', None)
        else:
            os.environ['TEMD_EDDY_MODE'] = old
    for q in range(3):
        assert nerr(out['split'][q].cpu().numpy(), out['fused'][q].cpu().numpy()) < 1e-11
