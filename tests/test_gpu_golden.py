"""GPU: the CUDA path against outputs of the unmodified reference (tests/golden/*.npz), through the
public API with DataArray inputs (tests/golden/_xarray_shim.DataArray stands in for xarray)."""
import glob
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
GOLDEN = sorted(glob.glob(os.path.join(HERE, 'golden', 'tem_*.npz')))
TOL = 1e-10
METHODS = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')
PROPS = ('ub', 'vb', 'thetab', 'wapb', 'upvpb', 'upwappb', 'vptpb', 'dub_dp', 'dthetab_dp', 'ubcoslat',
         'dubcoslat_dlat', 'psi', 'psicoslat', 'dpsicoslat_dlat', 'dpsi_dp', 'int_vbdp')


def nerr(x, ref):
    return float(np.abs(np.asarray(x) - ref).max() / max(np.abs(ref).max(), 1e-300))


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_tem_against_reference_outputs(path):
    from _xarray_shim import DataArray
    from pytemdiags_b200 import TEMDiagnostics
    g = np.load(path)
    dims = tuple(str(d) for d in g['dims'])
    coords = {'plev': g['plev_in'], 'time': g['time']}
    da = {n: DataArray(g['in_' + n], dims=dims, coords=coords, name=n) for n in ('ua', 'va', 'ta', 'wap')}
    lat = DataArray(g['lat'], dims=('ncol',), name='lat')
    kw = {} if float(g['zm_dlat']) == 1 else {'zm_dlat': int(g['zm_dlat'])}
    if 'in_q' in g:
        kw['q'] = DataArray(g['in_q'], dims=dims, coords=coords, name='q')
    tem = TEMDiagnostics(da['ua'], da['va'], da['ta'], da['wap'], lat, L=int(g['L']), debug_level=0, **kw)
    for m in METHODS:
        r = getattr(tem, m)()
        assert isinstance(r, DataArray) and r.dims == ('lat', 'plev', 'time') and r.name == m
        assert np.array_equal(r.coords['lat'], g['lat_zm'])
        assert np.allclose(r.coords['plev'] * 100, g['p'], rtol=0, atol=0)      # model top first
        assert nerr(r.values, g['ref_' + m]) < TOL, (m, nerr(r.values, g['ref_' + m]))
    for p_ in PROPS:
        assert nerr(getattr(tem, p_).values, g['ref_' + p_]) < TOL, p_
    if 'in_q' in g:     # tracer TEM (tem_diagnostics.py:801-991)
        for m in ('etfy', 'etfz', 'etdiv', 'qtendetfd', 'qtendvtem', 'qtendwtem'):
            r = getattr(tem, m)()
            assert r.dims == ('lat', 'plev', 'time') and r.name == m
            assert nerr(r.values, g['ref_' + m + '0']) < TOL, (m, nerr(r.values, g['ref_' + m + '0']))
        for p_ in ('qb', 'qpvpb', 'qpwappb', 'dqb_dp', 'qbcoslat', 'dqbcoslat_dlat'):
            assert nerr(getattr(tem, p_)[0].values, g['ref_' + p_ + '0']) < TOL, p_
    # the stand-alone averager on the same grid (sph_zonal_mean.py:285-296)
    A = da['ua'].transpose('ncol', 'plev', 'time')
    zm = tem.ZM.sph_zonal_mean(A)
    assert zm.dims == ('lat', 'plev', 'time') and zm.attrs['long_name'] == 'zonal mean of ua'
    assert nerr(zm.values, g['ref_zm_ua']) < TOL
    assert nerr(tem.ZM.sph_zonal_mean_native(A).values, g['ref_zmnative_ua']) < TOL
    assert nerr(tem.ZM.Y0inv, g['ref_Y0inv']) < TOL
