"""CPU: pins the NumPy oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by
tests/golden/make_golden.py) and against the reference's analytic known-answer checks."""
import glob
import os

import numpy as np
import pytest

import oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'tem_*.npz')))
TRACER = ('etfy', 'etfz', 'etdiv', 'qtendetfd', 'qtendvtem', 'qtendwtem', 'qb', 'qpvpb', 'qpwappb', 'dqb_dp',
          'qbcoslat', 'dqbcoslat_dlat')


def nerr(x, ref):
    return float(np.abs(np.asarray(x) - ref).max() / max(np.abs(ref).max(), 1e-300))


def load_case(path):
    g = np.load(path)
    dims = tuple(str(d) for d in g['dims'])
    to_nkt = [dims.index(d) for d in ('ncol', 'plev', 'time')]
    fields = {n: np.transpose(g['in_' + n], to_nkt) for n in ('ua', 'va', 'ta', 'wap', 'q') if 'in_' + n in g}
    return g, fields


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize('literal', [True, False])
def test_oracle_matches_reference(path, literal):
    g, f = load_case(path)
    out = oracle.tem_suite(f['ua'], f['va'], f['ta'], f['wap'], g['plev_in'], g['lat'], L=int(g['L']),
                           zm_dlat=float(g['zm_dlat']) if float(g['zm_dlat']) != 1 else 1, literal=literal, q=f.get('q'))
    assert np.array_equal(out['lat_zm'], g['lat_zm'])
    assert np.allclose(out['p'], g['p'], rtol=0, atol=0)
    tol = 1e-13 if literal else 1e-11
    for n in oracle.TEM_OUTPUTS + oracle.TEM_INTERMEDIATES:
        assert out[n].shape == g['ref_' + n].shape, n
        assert nerr(out[n], g['ref_' + n]) < tol, (n, nerr(out[n], g['ref_' + n]))
    if 'q' in f:
        for n in TRACER:
            assert nerr(out[n + '0'], g['ref_' + n + '0']) < tol, n


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_zonal_mean_and_matrices(path):
    g, f = load_case(path)
    Y0, Y0inv, Y0p = oracle.sph_matrices(g['lat'], g['lat_zm'], int(g['L']))
    assert nerr(Y0inv, g['ref_Y0inv']) < 1e-13
    flip = g['plev_in'][0] > g['plev_in'][-1]
    ua = f['ua']   # the fixture's stand-alone zonal mean used the input's own level order
    assert nerr(oracle.zonal_mean(ua, Y0p, Y0inv), g['ref_zm_ua']) < 1e-13
    assert nerr(oracle.zonal_mean(ua, Y0, Y0inv), g['ref_zmnative_ua']) < 1e-13
    # reference's logged sanity check, sph_zonal_mean.py:393-398
    P = Y0inv @ Y0
    assert abs(np.trace(P) - (int(g['L']) + 1)) < 1e-9 and abs(P.sum() - np.trace(P)) < 1e-8
    assert flip in (True, False)


def test_known_answers_analytic():
    """tests_sph_zonal_mean.py:331-347,465-475 re-derived on a synthetic pg2 grid."""
    import scipy.special as sp
    from pytemdiags_b200 import synthetic as syn
    lat, lon = syn.pg2_grid(15)
    lat_out = np.arange(-89.5, 90.5, 1)
    colat, colat_out, lonr = np.deg2rad(90 - lat), np.deg2rad(90 - lat_out), np.deg2rad(lon)
    Y0, Y0inv, Y0p = oracle.sph_matrices(lat, lat_out, 40)
    zm = lambda A: oracle.zonal_mean(A, Y0p, Y0inv)
    assert np.abs(zm(sp.sph_harm_y(2, 1, colat, lonr).real)).max() < 1e-3
    assert np.abs(zm(np.sin(lonr))).max() < 2e-2
    assert np.allclose(zm(sp.sph_harm_y(2, 0, colat, lonr).real), sp.sph_harm_y(2, 0, colat_out, 0).real, atol=1e-12)
    assert np.allclose(zm(np.deg2rad(lat) ** 2 + 1), np.deg2rad(lat_out) ** 2 + 1, rtol=2e-2)
    errs = []
    for L in (10, 20, 40):
        a, b, c = oracle.sph_matrices(lat, lat_out, L)
        errs.append(np.abs(oracle.zonal_mean(np.deg2rad(lat) ** 2 + 1, c, b) / (np.deg2rad(lat_out) ** 2 + 1) - 1).max())
    assert errs[0] > errs[1] > errs[2]          # error decreases with L (:462)


def test_recurrence_matches_scipy_and_mpmath():
    import mpmath
    x = np.cos(np.deg2rad(90 - np.array([-89.5, -60.0, -10.3, 0.0, 33.3, 75.0, 89.9])))
    R = oracle.sph_basis_recurrence(x, 700)
    S = oracle.sph_basis(np.array([-89.5, -60.0, -10.3, 0.0, 33.3, 75.0, 89.9]), 700)
    assert np.abs(R[:, :646] - S[:, :646]).max() < 5e-11
    assert np.array_equal(R[:, 646:], S[:, 646:])       # beyond SciPy's range the oracle IS the recurrence
    mpmath.mp.dps = 40
    for l in (200, 700):
        for xi, ri in zip(x[1:4], R[1:4, l]):
            exact = mpmath.sqrt((2 * l + 1) / (4 * mpmath.pi)) * mpmath.legendre(l, mpmath.mpf(float(xi)))
            assert abs(float(exact) - ri) < 2e-10


def test_p_integral_is_cumulative_trapezoid():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((5, 9, 3))
    p = np.sort(rng.uniform(1, 1000, 9)) * 100
    from oracle.tem_oracle import _p_integral
    ref = _p_integral(A, p)
    cum = np.concatenate([np.zeros((5, 1, 3)), np.cumsum(np.diff(p)[None, :, None] * (A[:, 1:] + A[:, :-1]) / 2.0, axis=1)], axis=1)
    assert np.allclose(ref, cum, rtol=1e-14, atol=1e-9)


def test_basis_latitude_derivative_against_mpmath():
    """Optional Legendre-space derivative (north_star): d/dphi [sqrt((2l+1)/4pi) P_l(sin phi)] from the differentiated
    recurrence vs mpmath at 50 digits (closed form l (P_{l-1} - x P_l) / cos(phi) away from the poles, 0 at the poles),
    and vs a centred finite difference of the SciPy basis."""
    import mpmath as mp
    mp.mp.dps = 50
    lat = np.array([-90.0, -89.5, -60.0, -12.25, 0.0, 33.0, 75.5, 89.99, 90.0])
    L = 120
    D = oracle.sph_basis_dlat(lat, L)
    assert D.shape == (lat.shape[0], L + 1) and np.all(D[:, 0] == 0.0)
    for i, la in enumerate(lat):
        phi = mp.mpf(la) * mp.pi / 180
        x, c = mp.sin(phi), mp.cos(phi)
        for l in (1, 2, 3, 17, 64, 120):
            nl = mp.sqrt((2 * l + 1) / (4 * mp.pi))
            if abs(la) == 90.0:
                ref = mp.mpf(0)
            else:
                ref = nl * l * (mp.legendre(l - 1, x) - x * mp.legendre(l, x)) / c
            scale = float(nl) * l * (l + 1) / 2          # max |dY_l/dphi| is of this order
            assert abs(D[i, l] - float(ref)) < 1e-12 * scale, (la, l, D[i, l], float(ref))
    h = 1e-5
    lat2 = np.array([-70.0, -20.0, 5.0, 48.0])
    fd = (oracle.sph_basis(lat2 + np.rad2deg(h), 30) - oracle.sph_basis(lat2 - np.rad2deg(h), 30)) / (2 * h)
    assert np.abs(oracle.sph_basis_dlat(lat2, 30) - fd).max() < 1e-6
