"""CPU: host-side logic of the drop-in classes (argument shapes, validation, coordinates) with the GPU
engine stubbed out.  No arithmetic on field data happens here."""
import numpy as np
import pytest

import pytemdiags_b200.tem as tem_mod
from pytemdiags_b200 import synthetic as syn


class _StubAverager:
    def __init__(self, lat, lat_out, L, **kw):
        self.lat, self.lat_out, self.L, self.kw = lat, lat_out, L, kw
        self._engine = None

    def sph_compute_matrices(self, **kw):
        pass

    def sph_zonal_mean(self, A):
        raise NotImplementedError


@pytest.fixture
def stubbed(monkeypatch):
    monkeypatch.setattr(tem_mod, 'sph_zonal_averager', _StubAverager)
    monkeypatch.setattr(tem_mod.TEMDiagnostics, '_compute_all', lambda self: None)
    return tem_mod.TEMDiagnostics


def _fields(N=96, K=5, T=3):
    lat, lon = syn.pg2_grid(2)
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=0)
    return lat, plev, f


def test_both_call_shapes(stubbed):
    lat, plev, f = _fields()
    a = stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, dims=('time', 'lev', 'ncol'), debug_level=0)      # README shape
    b = stubbed(f['ua'], f['va'], f['ta'], f['wap'], lat, p=plev, dims=('time', 'lev', 'ncol'), debug_level=0)    # code shape
    c = stubbed(f['ua'], f['va'], f['ta'], f['wap'], lat_native=lat, p=plev, L=30, dims=('time', 'lev', 'ncol'), debug_level=0)
    for t in (a, b, c):
        assert t.NCOL == 96 and t.NLEV == 5 and t.NT == 3
        assert np.array_equal(t.plev, plev) and np.allclose(t.p, plev * 100)
        assert t.ZM_N == 180 and t.lat[0] == -89.5 and t.lat[-1] == 89.5
        assert t.f.shape == (180, 1, 1) and np.allclose(t.coslat, np.cos(np.deg2rad(t.lat)))
    assert a.L == 50 and c.L == 30


def test_pressure_conventions(stubbed):
    lat, plev, f = _fields()
    kw = dict(dims=('time', 'lev', 'ncol'), debug_level=0)
    t = stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev[::-1].copy(), lat, **kw)
    assert t._flip_lev and np.array_equal(t.plev, plev)          # outputs are always model-top-first (:372-382)
    t = stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev * 100, lat, p_units='Pa', **kw)
    assert np.allclose(t.plev, plev)
    pfield = np.broadcast_to((plev * 100)[None, :, None], f['ua'].shape).copy()        # gridpoint pressure [Pa]
    t = stubbed(f['ua'], f['va'], f['ta'], f['wap'], pfield, lat, **kw)
    assert np.allclose(t.plev, plev)
    pfield[0, 1, 3] *= 1.01
    with pytest.raises(RuntimeError, match='varies on a level'):
        stubbed(f['ua'], f['va'], f['ta'], f['wap'], pfield, lat, **kw)
    with pytest.raises(RuntimeError, match='pressure levels are required'):
        stubbed(f['ua'], f['va'], f['ta'], f['wap'], lat, **kw)


def test_zm_grid_options(stubbed):
    lat, plev, f = _fields()
    kw = dict(dims=('time', 'lev', 'ncol'), debug_level=0)
    t = stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, zm_dlat=2, **kw)
    assert t.ZM_N == 90 and t.lat[0] == -89.0
    t = stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, zm_pole_points=True, **kw)
    assert t.ZM_N == 181 and t.lat[0] == -90 and t.lat[-1] == 90
    with pytest.raises(AssertionError):
        stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, zm_dlat=7, **kw)


def test_validation_errors(stubbed):
    lat, plev, f = _fields()
    kw = dict(dims=('time', 'lev', 'ncol'), debug_level=0)
    with pytest.raises(RuntimeError, match='must match'):
        stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat[:-1], **kw)
    with pytest.raises(RuntimeError, match='does not contain dim'):
        stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, dims=('time', 'lev', 'cells'), debug_level=0)
    with pytest.raises(RuntimeError, match='different shapes'):
        stubbed(f['ua'], f['va'][:2], f['ta'], f['wap'], plev, lat, **kw)
    with pytest.raises(TypeError):
        stubbed(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, bogus=1, **kw)
    with pytest.raises(RuntimeError, match='has 4 dims|does not match'):
        stubbed(f['ua'][None], f['va'][None], f['ta'][None], f['wap'][None], plev, lat, **kw)


def test_two_dimensional_inputs_get_a_time_axis(stubbed):
    """The reference intends T=1 for (ncol, plev) inputs (tem_diagnostics.py:332-335, broken there)."""
    lat, plev, f = _fields(T=1)
    g = {k: v[0].T.copy() for k, v in f.items()}      # (ncol, plev)
    t = stubbed(g['ua'], g['va'], g['ta'], g['wap'], lat, p=plev, debug_level=0)
    assert (t.NCOL, t.NLEV, t.NT) == (96, 5, 1)


def test_dataarray_inputs_any_dim_order(stubbed):
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from _xarray_shim import DataArray
    lat, plev, f = _fields()
    time = np.arange(3) * 6.0
    da = {k: DataArray(np.ascontiguousarray(v.transpose(1, 2, 0)), dims=('plev', 'ncol', 'time'),
                       coords={'plev': plev, 'time': time}, name=k) for k, v in f.items()}
    t = stubbed(da['ua'], da['va'], da['ta'], da['wap'], DataArray(lat, dims=('ncol',)), debug_level=0)
    assert (t.NCOL, t.NLEV, t.NT) == (96, 5, 3) and np.array_equal(t.time, time) and np.array_equal(t.plev, plev)
    t = stubbed(da['ua'], da['va'], da['ta'], da['wap'], DataArray(lat, dims=('ncol',)), q=[da['ua'], da['va']], debug_level=0)
    assert t.ntrac == 2
    with pytest.raises(RuntimeError, match='qi must be passed'):
        t._qi(None, 'etfy')


def test_format_latlon_data_column_order():
    """lat-major ravel, ncol = ilat*NLON + ilon (tem_util.py:331); matches synthetic.latlon_grid (config 4's layout)."""
    from pytemdiags_b200.util import format_latlon_data
    lat = np.linspace(-90, 90, 5)
    lon = np.arange(8) * 45.0
    A = np.arange(3 * 5 * 8, dtype=np.float64).reshape(3, 5, 8)          # (time, lat, lon)
    A2, la, lo = format_latlon_data(A, lat, lon)
    assert A2.shape == (3, 40) and A2[1, 2 * 8 + 3] == A[1, 2, 3]
    glat, glon = syn.latlon_grid(5, 8)
    assert np.array_equal(la, glat) and np.array_equal(lo, glon)
    B2, _, _ = format_latlon_data(A.transpose(2, 0, 1), lat, lon, lat_axis=2, lon_axis=0)
    assert np.array_equal(B2, A2)
    with pytest.raises(RuntimeError):
        format_latlon_data(A, lat[:-1], lon)


def test_format_latlon_data_reference_signature():
    """Dataset in, Dataset out, as tem_util.py:247-342 (a mapping name -> (dims, array) stands in for xarray.Dataset):
    (lat, lon) stacked lat-major into a leading 'ncol', lat / lon kept as ('ncol',) variables, bounds variables added."""
    from pytemdiags_b200.util import format_latlon_data
    import PyTEMDiags
    lat = np.linspace(-90, 90, 5)
    lon = np.arange(8) * 45.0
    A = np.arange(3 * 5 * 8, dtype=np.float64).reshape(3, 5, 8)
    ds = {'ua': (('time', 'lat', 'lon'), A), 'ps': (('lat', 'lon'), A[0]), 'hyam': (('lev',), np.arange(4.0)),
          'lat': (('lat',), lat), 'lon': (('lon',), lon)}
    out = format_latlon_data(ds)
    assert out['ua'][0] == ('ncol', 'time') and out['ua'][1].shape == (40, 3)
    assert out['ua'][1][2 * 8 + 3, 1] == A[1, 2, 3]                     # ncol = ilat * NLON + ilon (:331)
    assert out['ps'][0] == ('ncol',) and np.array_equal(out['ps'][1], A[0].ravel())
    assert out['hyam'][0] == ('lev',)                                     # variables without lat/lon pass through
    glat, glon = syn.latlon_grid(5, 8)
    assert out['lat'][0] == ('ncol',) and np.array_equal(out['lat'][1], glat) and np.array_equal(out['lon'][1], glon)
    # bounds at the midpoints between neighbours (:309-323), stacked like every other (lat)/(lon) variable
    assert out['lat_bnds'][0] == ('ncol', 'nbnd') and out['lat_bnds'][1].shape == (40, 2)
    assert np.allclose(out['lat_bnds'][1][8], [-67.5, -22.5]) and np.allclose(out['lon_bnds'][1][3], [112.5, 157.5])
    # custom names, existing bounds with a wrong bound-dimension name -> the reference's RuntimeError
    ds2 = {'t': (('latitude', 'longitude'), A[0]), 'latitude': (('latitude',), lat), 'longitude': (('longitude',), lon),
           'lat_bnds': (('latitude', 'bnds'), np.zeros((5, 2)))}
    with pytest.raises(RuntimeError, match='does not have dimension'):
        format_latlon_data(ds2, 'latitude', 'longitude')
    out2 = format_latlon_data(ds2, lat_name='latitude', lon_name='longitude', bnddim_name='bnds')
    assert out2['t'][0] == ('ncol',)
    assert hasattr(PyTEMDiags, 'tem_util') and PyTEMDiags.tem_util.format_latlon_data is format_latlon_data


def test_slab_schedule_and_rank_check():
    """Host helpers of the streaming path: the tapered slab schedule covers the record exactly once, in order; the
    early rank-deficiency check counts distinct latitudes (default L=50 on a coarse lat-lon grid)."""
    from pytemdiags_b200.engine import _check_rank
    for T, ts in ((16, 2), (16, 8), (17, 8), (5, 8), (1, 1), (3, 2), (365, 73), (96, 12)):
        s = tem_mod._slab_schedule(T, ts)
        assert s[0][0] == 0 and s[-1][1] == T and all(a[1] == b[0] for a, b in zip(s, s[1:]))
        assert all(0 < t1 - t0 <= ts for t0, t1 in s)
        if T > ts:
            assert s[-1][1] - s[-1][0] == 1                 # the tail after the last copy is one step
    lat, _ = syn.latlon_grid(24, 48)
    with pytest.raises(RuntimeError, match='only 24 distinct latitudes'):
        _check_rank(lat, 50)
    _check_rank(lat, 23)                                     # L + 1 = 24 <= 24 distinct latitudes: fine
    _check_rank(lat[:5], 6)                                  # tiny problems (L + 1 <= 8) are left to the factorisation
