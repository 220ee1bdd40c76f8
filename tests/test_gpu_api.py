"""Public-API parity (TEMDiagnostics / sph_zonal_averager) against the CPU oracle on a B200."""
import numpy as np
import pytest

import oracle
from pytemdiags_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-10


def nerr(x, ref):
    return float(np.abs(np.asarray(x) - ref).max() / max(np.abs(ref).max(), 1e-300))


def _case(ne, K, T, seed, plev=None, latlon=None):
    if latlon:
        lat, lon = syn.latlon_grid(*latlon)
    else:
        lat, lon = syn.pg2_grid(ne)
    plev = syn.default_plev(K) if plev is None else plev
    f = syn.synth_fields(lat, lon, plev, T, seed=seed)
    return lat, lon, plev, f


def _ref(f, plev, lat, L, **kw):
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0))
    return oracle.tem_suite(tr(f['ua']), tr(f['va']), tr(f['ta']), tr(f['wap']), plev, lat, L=L, **kw)


ALL = oracle.TEM_OUTPUTS + oracle.TEM_INTERMEDIATES


def _check(tem, ref, tol=TOL):
    for n in ALL:
        got = getattr(tem, n)
        got = got() if callable(got) else got
        assert got.shape == ref[n].shape, n
        assert nerr(got, ref[n]) < tol, (n, nerr(got, ref[n]))


@pytest.mark.parametrize('ne,K,T,L', [(4, 6, 2, 10), (8, 12, 3, 50), (16, 10, 2, 100)])
def test_tem_suite_time_lev_ncol(ne, K, T, L):
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon, plev, f = _case(ne, K, T, seed=1)
    tem = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0)
    _check(tem, _ref(f, plev, lat, L))


def test_tem_suite_reference_layout_and_reversed_plev():
    """(ncol, plev, time) inputs with plev descending: outputs must come back model-top-first
    (tem_diagnostics.py:372-382)."""
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon, plev, f = _case(6, 9, 3, seed=2)
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0)[:, ::-1, :])
    tem = TEMDiagnostics(tr(f['ua']), tr(f['va']), tr(f['ta']), tr(f['wap']), lat, p=plev[::-1].copy(), L=25, debug_level=0)
    ref = _ref(f, plev, lat, 25)
    _check(tem, ref)
    assert np.array_equal(tem.plev, plev)


def test_tem_suite_slabbed_and_float32():
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon, plev, f = _case(6, 7, 5, seed=3)
    f32 = {k: v.astype(np.float32) for k, v in f.items()}
    N = lat.shape[0]
    tem = TEMDiagnostics(f32['ua'], f32['va'], f32['ta'], f32['wap'], plev, lat, L=25, dims=('time', 'lev', 'ncol'),
                         debug_level=0, slab_bytes=2 * 4 * 8 * 7 * N)   # two time steps per slab
    ref = _ref({k: v.astype(np.float64) for k, v in f32.items()}, plev, lat, 25)
    assert tem.vtem().dtype == np.float32
    for n in oracle.TEM_OUTPUTS:
        assert nerr(getattr(tem, n)(), ref[n]) < 1e-6, n      # float32 output rounding


def test_tem_suite_torch_cuda_inputs_and_pole_points():
    import torch
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon, plev, f = _case(8, 8, 2, seed=4)
    dv = {k: torch.as_tensor(v).cuda() for k, v in f.items()}
    tem = TEMDiagnostics(dv['ua'], dv['va'], dv['ta'], dv['wap'], plev, lat, L=30, dims=('time', 'lev', 'ncol'),
                         debug_level=0, zm_dlat=2, zm_pole_points=False)
    ref = _ref(f, plev, lat, 30, zm_dlat=2)
    for n in oracle.TEM_OUTPUTS:
        got = getattr(tem, n)()
        assert isinstance(got, torch.Tensor) and got.is_cuda
        assert nerr(got.cpu().numpy(), ref[n]) < TOL, n


def test_tem_latlon_grid():
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon, plev, f = _case(None, 6, 2, seed=5, latlon=(45, 90))
    tem = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=20, dims=('time', 'lev', 'ncol'), debug_level=0)
    _check(tem, _ref(f, plev, lat, 20))


def test_zonal_averager_known_answers():
    """Analytic checks of reference tests_sph_zonal_mean.py:331-347,465-475 on a synthetic pg2 grid."""
    import scipy.special as sp
    from pytemdiags_b200 import sph_zonal_averager
    lat, lon = syn.pg2_grid(15)
    lat_out = np.arange(-89.5, 90.5, 1)
    colat, colat_out = np.deg2rad(90 - lat), np.deg2rad(90 - lat_out)
    lonr = np.deg2rad(lon)
    ZM = sph_zonal_averager(lat, lat_out, 40, overwrite=True)
    assert ZM.Y0 is None
    with pytest.raises(RuntimeError):
        ZM.sph_zonal_mean(np.zeros(lat.shape[0]))
    ZM.sph_compute_matrices()
    y21 = sp.sph_harm_y(2, 1, colat, lonr).real
    y20 = sp.sph_harm_y(2, 0, colat, lonr).real
    f1 = np.sin(lonr)
    f2 = np.deg2rad(lat) ** 2 + 1
    assert np.abs(ZM.sph_zonal_mean(y21)).max() < 1e-3
    assert np.abs(ZM.sph_zonal_mean(f1)).max() < 2e-2
    assert np.allclose(ZM.sph_zonal_mean(y20), sp.sph_harm_y(2, 0, colat_out, 0).real, atol=1e-12)
    assert np.allclose(ZM.sph_zonal_mean(f2), np.deg2rad(lat_out) ** 2 + 1, rtol=2e-2)
    # zonally symmetric input is reproduced on the native grid (tests_sph_zonal_mean.py:152-204)
    assert np.allclose(ZM.sph_zonal_mean_native(y20), y20, atol=1e-12)
    # multi-dimensional input, reference layout (ncol, lev, time), and the oracle
    A = np.random.default_rng(0).standard_normal((lat.shape[0], 3, 4))
    Y0, Y0inv, Y0p = oracle.sph_matrices(lat, lat_out, 40)
    assert nerr(ZM.sph_zonal_mean(A), oracle.zonal_mean(A, Y0p, Y0inv)) < TOL
    assert nerr(ZM.sph_zonal_mean_native(A), oracle.zonal_mean(A, Y0, Y0inv)) < TOL
    assert nerr(ZM.Y0inv, Y0inv) < TOL and nerr(ZM.Y0, Y0) < 1e-12 and nerr(ZM.Y0p, Y0p) < 1e-12


def test_error_behaviour():
    from pytemdiags_b200 import TEMDiagnostics, sph_zonal_averager
    lat, lon, plev, f = _case(4, 5, 2, seed=6)
    bad = f['ua'].copy()
    bad[1, 2, 3] = np.nan
    with pytest.raises(RuntimeError, match='nans'):
        TEMDiagnostics(bad, f['va'], f['ta'], f['wap'], plev, lat, L=10, dims=('time', 'lev', 'ncol'), debug_level=0)
    with pytest.raises(RuntimeError, match='must match'):
        TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat[:-2], L=10, dims=('time', 'lev', 'ncol'), debug_level=0)
    with pytest.raises(AssertionError):
        TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=10, dims=('time', 'lev', 'ncol'), debug_level=0, zm_dlat=7)
    with pytest.raises(RuntimeError, match='rank-deficient'):
        sph_zonal_averager(lat, np.arange(-89.5, 90, 1.0), 300).sph_compute_matrices()
    # the reference's default L=50 on a lat-lon grid with fewer than 51 latitudes is rank-deficient (gelsd
    # returns a minimum-norm fit there); this build says so, with the remedy, for the dense and the dedup engine
    lat_ll, _ = syn.latlon_grid(24, 48)
    for kw in ({}, {'dedup': True}):
        with pytest.raises(RuntimeError, match='only 24 distinct latitudes.*lower L to at most 23'):
            sph_zonal_averager(lat_ll, np.arange(-89.5, 90, 1.0), 50, **kw)
    ZM = sph_zonal_averager(lat, np.arange(-89.5, 90, 1.0), 8)
    ZM.sph_compute_matrices()
    with pytest.raises(RuntimeError, match='length'):
        ZM.sph_zonal_mean(np.zeros(lat.shape[0] + 1))
    x = np.zeros(lat.shape[0]); x[5] = np.nan
    with pytest.raises(RuntimeError, match='nans'):
        ZM.sph_zonal_mean(x)
    # the reference screens NaN only (sph_zonal_mean.py:219): an infinity is reported as what it is, not as "nans"
    x[5] = np.inf
    with pytest.raises(RuntimeError, match='infinite') as ei:
        ZM.sph_zonal_mean(x)
    assert 'has nans' not in str(ei.value)
    badi = f['wap'].copy()
    badi[0, 1, 7] = -np.inf
    with pytest.raises(RuntimeError, match='infinite'):
        TEMDiagnostics(f['ua'], f['va'], f['ta'], badi, plev, lat, L=10, dims=('time', 'lev', 'ncol'), debug_level=0)
    badi[1, 0, 0] = np.nan          # NaN wins over infinity, as in the reference
    with pytest.raises(RuntimeError, match='nans'):
        TEMDiagnostics(f['ua'], f['va'], f['ta'], badi, plev, lat, L=10, dims=('time', 'lev', 'ncol'), debug_level=0)


def test_zonal_averager_weights_path():
    """Deprecated quadrature inverse Y0inv = Y0^T diag(4 pi w) (sph_zonal_mean.py:180-181,383-386)."""
    from pytemdiags_b200 import sph_zonal_averager
    lat, lon = syn.pg2_grid(10)
    lat_out = np.arange(-89.5, 90.5, 1)
    w = np.cos(np.deg2rad(lat)); w /= w.sum()
    w0 = w.copy()
    ZM = sph_zonal_averager(lat, lat_out, 20, weights=w)
    ZM.sph_compute_matrices()
    assert np.array_equal(w, w0)
    Y0 = oracle.sph_basis(lat, 20)
    Y0p = oracle.sph_basis(lat_out, 20)
    Y0inv = np.matmul(Y0.T, np.diag(4 * np.pi * w))
    A = np.random.default_rng(1).standard_normal((lat.shape[0], 4, 2))
    assert nerr(ZM.sph_zonal_mean(A), oracle.zonal_mean(A, Y0p, Y0inv)) < TOL
    assert nerr(ZM.sph_zonal_mean_native(A), oracle.zonal_mean(A, Y0, Y0inv)) < TOL
    assert nerr(ZM.Y0inv, Y0inv) < 1e-13 and nerr(ZM.Y0, Y0) < 1e-12


def test_native_grid_properties_on_demand():
    """up, vp, thetap, wapp, upvp, upwapp, vptp (tem_diagnostics.py:420-433) are rebuilt on request."""
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon, plev, f = _case(6, 7, 3, seed=8)
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0)[:, ::-1, :])       # (ncol, plev, time), plev descending
    tem = TEMDiagnostics(tr(f['ua']), tr(f['va']), tr(f['ta']), tr(f['wap']), lat, p=plev[::-1].copy(), L=25, debug_level=0)
    ref = _ref(f, plev, lat, 25)
    for n in ('up', 'vp', 'thetap', 'wapp'):
        got = getattr(tem, n)
        assert got.shape == ref[n].shape and nerr(got, ref[n]) < TOL, n
    assert nerr(tem.upvp, ref['up'] * ref['vp']) < TOL
    assert nerr(tem.upwapp, ref['up'] * ref['wapp']) < TOL
    assert nerr(tem.vptp, ref['vp'] * ref['thetap']) < TOL
    assert nerr(tem.theta, ref['theta']) < 1e-14


def test_odd_column_count_and_pole_points():
    """N odd (the TMA path needs an even leading dimension: the host pads) and zm_pole_points=True
    (M = 181 odd; 1/(a cos(lat)) is ~1e9 at the poles, same arithmetic as the reference, tem_diagnostics.py:391-394)."""
    from pytemdiags_b200 import TEMDiagnostics, sph_zonal_averager
    lat, lon = syn.latlon_grid(15, 7, poles=False)
    assert lat.shape[0] % 2 == 1
    K, T, L = 6, 2, 8
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=9)
    tem = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'),
                         debug_level=0, zm_pole_points=True)
    ref = _ref(f, plev, lat, L, zm_pole_points=True)
    assert tem.ZM_N == 181
    for n in ALL:
        got = getattr(tem, n)
        got = got() if callable(got) else got
        assert got.shape == ref[n].shape and nerr(got, ref[n]) < TOL, (n, nerr(got, ref[n]))
    ZM = sph_zonal_averager(lat, np.arange(-89.5, 90, 1.0), L)
    ZM.sph_compute_matrices()
    A = np.random.default_rng(3).standard_normal((lat.shape[0], 5))
    Y0, Y0inv, Y0p = oracle.sph_matrices(lat, np.arange(-89.5, 90, 1.0), L)
    assert nerr(ZM.sph_zonal_mean(A), oracle.zonal_mean(A, Y0p, Y0inv)) < TOL
    assert nerr(ZM.sph_zonal_mean_native(A), oracle.zonal_mean(A, Y0, Y0inv)) < TOL


def test_large_L_staged_path():
    """L + 1 > 408 exceeds the fused kernel's shared-memory budget: the staged GPU path must give the same answers."""
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon = syn.pg2_grid(72)            # 124,416 columns: cond(Y0) = 6 at L = 420
    K, T, L = 5, 1, 420
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=10)
    tem = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'), debug_level=0)
    assert tem.ZM._engine.lpad > 408
    mats = oracle.sph_matrices(lat, oracle.zm_latitudes(1), L, method='normal', basis='recurrence')
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0))
    ref = oracle.tem_suite(tr(f['ua']), tr(f['va']), tr(f['ta']), tr(f['wap']), plev, lat, L=L, literal=False, matrices=mats)
    for n in ALL:
        got = getattr(tem, n)
        got = got() if callable(got) else got
        assert nerr(got, ref[n]) < TOL, (n, nerr(got, ref[n]))


def test_ill_conditioned_basis_uses_shifted_cholesky():
    """cond(Y0) = 3e9 (L close to the grid's resolution): plain CholeskyQR breaks down, the shifted first pass
    must rescue it; the projector still agrees with pinv to ~cond * eps."""
    from pytemdiags_b200 import sph_zonal_averager
    lat, lon = syn.pg2_grid(8)
    lat_out = np.arange(-89.5, 90, 1.0)
    L = 100
    ZM = sph_zonal_averager(lat, lat_out, L, debug=True)
    ZM.sph_compute_matrices()
    assert abs(ZM._engine.sanity[0] - (L + 1)) < 1e-8 and abs(ZM._engine.sanity[1]) < 1e-7     # Q^T Q = I
    A = syn.synth_fields(lat, lon, syn.default_plev(4), 1, seed=5, fields=('ua',))['ua'][0].T.copy()   # (N, 4)
    Y0, Y0inv, Y0p = oracle.sph_matrices(lat, lat_out, L, method='pinv')
    assert nerr(ZM.sph_zonal_mean_native(A), oracle.zonal_mean(A, Y0, Y0inv)) < 1e-6
    assert nerr(ZM.sph_zonal_mean(A), oracle.zonal_mean(A, Y0p, Y0inv)) < 1e-4


def test_many_rows_and_fine_zm_grid():
    """rows = time*lev = 72,000 (> 65,535: no 16-bit grid-dimension limits anywhere) and zm_dlat = 0.5 (M = 360)."""
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon = syn.pg2_grid(4)
    K, T, L = 72, 1000, 10
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=11)
    tem = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=L, dims=('time', 'lev', 'ncol'),
                         debug_level=0, zm_dlat=0.5)
    ref = _ref(f, plev, lat, L, zm_dlat=0.5)
    assert tem.ZM_N == 360
    for n in oracle.TEM_OUTPUTS:
        assert nerr(getattr(tem, n)(), ref[n]) < TOL, n
    assert nerr(tem.up, ref['up']) < TOL          # on-demand native eddy over 72,000 rows


def test_to_netcdf_roundtrip(tmp_path):
    """to_netcdf / q_to_netcdf (tem_diagnostics.py:995-1103): same file names and variables, NetCDF-3 via SciPy."""
    from scipy.io import netcdf_file
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon = syn.pg2_grid(4)
    K, T, L = 5, 2, 10
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=12, fields=('ua', 'va', 'ta', 'wap', 'q'))
    tem = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, q=f['q'], L=L, dims=('time', 'lev', 'ncol'),
                         debug_level=0, grid_name='ne4pg2')
    path = tem.to_netcdf(loc=str(tmp_path), prefix='x', include_attrs=True)
    assert path.endswith('x_TEM_ne4pg2_1.0deg_L10.nc') and tem.out_file == path
    with netcdf_file(path, 'r', mmap=False) as nc:
        for n in oracle.TEM_OUTPUTS:
            assert nc.variables[n].dimensions == ('lat', 'plev', 'time')
            assert np.array_equal(nc.variables[n][:], getattr(tem, n)())
        assert nc.variables['up'].dimensions == ('ncol', 'plev', 'time')
        assert np.array_equal(nc.variables['lat'][:], tem.lat)
    qpaths = tem.q_to_netcdf(loc=str(tmp_path))
    assert qpaths[0].endswith('TEM_ne4pg2_1.0deg_L10_TRACER-q0.nc')
    with netcdf_file(qpaths[0], 'r', mmap=False) as nc:
        assert np.array_equal(nc.variables['etdiv'][:], tem.etdiv())


def test_spectral_latitude_derivative():
    """Optional extra (north_star): Legendre-space d/dphi of the zonal mean.  dY0p against the mpmath-pinned oracle;
    sph_zonal_mean_dlat against dY0p (Y0inv A) and, for a band-limited field, against the analytic derivative."""
    from pytemdiags_b200 import sph_zonal_averager
    lat, lon = syn.pg2_grid(12)
    lat_out = np.concatenate([[-90.0], np.arange(-89.5, 90, 1.0), [90.0]])        # poles included
    L = 30
    for kw in ({}, {'dedup': True}):
        ZM = sph_zonal_averager(lat, lat_out, L, **kw)
        ZM.sph_compute_matrices()
        dY0p = oracle.sph_basis_dlat(lat_out, L)
        assert nerr(ZM.dY0p, dY0p) < 1e-13
        A = np.random.default_rng(2).standard_normal((lat.shape[0], 3, 2))
        Y0, Y0inv, Y0p = oracle.sph_matrices(lat, lat_out, L)
        ref = (dY0p @ (Y0inv @ A.reshape(lat.shape[0], -1))).reshape((lat_out.shape[0], 3, 2))
        got = ZM.sph_zonal_mean_dlat(A)
        assert got.shape == ref.shape and nerr(got, ref) < TOL
        # f = sin(phi)^2 is exactly representable (degrees 0 and 2): d/dphi = sin(2 phi)
        f = np.sin(np.deg2rad(lat)) ** 2
        assert np.abs(ZM.sph_zonal_mean_dlat(f) - np.sin(2 * np.deg2rad(lat_out))).max() < 1e-11
    w = np.cos(np.deg2rad(lat)); w /= w.sum()
    ZW = sph_zonal_averager(lat, lat_out, 8, weights=w)                           # quadrature inverse: raw derivative basis
    ZW.sph_compute_matrices()
    Y0 = oracle.sph_basis(lat, 8)
    A = np.random.default_rng(3).standard_normal((lat.shape[0], 2))
    ref = oracle.sph_basis_dlat(lat_out, 8) @ ((Y0.T * (4 * np.pi * w)) @ A)
    assert nerr(ZW.sph_zonal_mean_dlat(A), ref) < TOL


@pytest.mark.parametrize('L', [24, 120])        # fused tracer-pair kernel / split path in pair mode
def test_several_tracers_pair_kernel(L):
    """Tracers are processed two per launch (temd_tracer_flux_project: q1'v', q1'omega', q2'v', q2'omega' with v' and
    omega' synthesised once) plus the TEM kernel for an odd one out.  Every tracer must give what it gives alone
    (the single-tracer path is pinned against the unmodified reference by the golden fixture)."""
    from pytemdiags_b200 import TEMDiagnostics
    lat, lon = syn.pg2_grid(10)
    K, T = 6, 3
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=31, fields=('ua', 'va', 'ta', 'wap', 'q'))
    rng = np.random.default_rng(8)
    qs = [f['q'], f['q'] * 0.5 + 1e-4 * rng.standard_normal(f['q'].shape), 3e-4 * np.abs(f['va']) + 1e-5]
    kw = dict(L=L, dims=('time', 'lev', 'ncol'), debug_level=0)
    both = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, q=qs, **kw)
    assert both.ntrac == 3
    for i, q in enumerate(qs):
        alone = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, q=q, **kw)
        for m in ('etfy', 'etfz', 'etdiv', 'qtendetfd', 'qtendvtem', 'qtendwtem'):
            e = nerr(getattr(both, m)(i), getattr(alone, m)(0))
            assert e < 1e-11, (i, m, e)
        for p_ in ('qb', 'qpvpb', 'qpwappb', 'dqb_dp'):
            assert nerr(getattr(both, p_)[i], getattr(alone, p_)[0]) < 1e-11, (i, p_)
    with pytest.raises(RuntimeError, match='qi must be passed'):
        both.etfy()


def test_plan_is_reentrant_per_stream():
    """one split-K workspace per plan made a plan unusable from two streams.  Workspaces are now kept per
    (plan, stream): the same engine driven from two CUDA streams at once gives the serial results bit for bit."""
    import torch
    from pytemdiags_b200.engine import Engine
    lat, lon = syn.pg2_grid(16)
    eng = Engine(lat, np.arange(-89.5, 90, 1.0), 60).build_basis()
    rng = np.random.default_rng(0)
    xs = [torch.as_tensor(rng.standard_normal((96, lat.shape[0]))).cuda() for _ in range(2)]
    ref = [eng.project([x]) for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [None, None]
    for rep in range(5):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                got[i] = eng.project([xs[i]])
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(got[i], ref[i])


def test_pageable_staging_is_capped_and_releasable():
    """the pinned staging buffers for pageable inputs used to live forever on the cached engine."""
    import pytemdiags_b200
    from pytemdiags_b200 import TEMDiagnostics, tem as tem_mod
    lat, lon = syn.pg2_grid(12)
    K, T = 40, 10                                   # 11 MB per field: above the 8 MB staging threshold
    plev = syn.default_plev(K)
    f = syn.synth_fields(lat, lon, plev, T, seed=13)
    a = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=20, dims=('time', 'lev', 'ncol'), debug_level=0)
    assert len(tem_mod._HOST_STAGING) > 0
    pytemdiags_b200.release_host_staging()
    assert len(tem_mod._HOST_STAGING) == 0
    b = TEMDiagnostics(f['ua'], f['va'], f['ta'], f['wap'], plev, lat, L=20, dims=('time', 'lev', 'ncol'), debug_level=0)
    assert np.array_equal(a.epfy(), b.epfy())
    pytemdiags_b200.release_host_staging()
