"""Float64 NumPy/SciPy restatement of the PyTEMDiags hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it restates (paths relative to /root/reference).
No xarray: arrays are plain ndarrays in the reference's internal layout (ncol, plev, time).
Nothing here is imported by the product package `pytemdiags_b200`.
"""
import numpy as np
import scipy.linalg
import scipy.special

# PyTEMDiags/constants.py:6-14, verbatim (note the truncated pi used by psitem)
CONSTANTS = dict(
    P0=101325, R=287.058, Cp=1004.64, g0=9.80665, a=6.37123e6, Om=7.29212e-5,
    k=287.058 / 1004.64, H=7 * 1e3, pi=3.14159,
)

TEM_OUTPUTS = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv',
               'utendepfd', 'utendvtem', 'utendwtem')
TEM_INTERMEDIATES = ('ub', 'vb', 'thetab', 'wapb', 'upvpb', 'upwappb', 'vptpb', 'dub_dp',
                     'dthetab_dp', 'ubcoslat', 'dubcoslat_dlat', 'psi', 'psicoslat',
                     'dpsicoslat_dlat', 'dpsi_dp', 'int_vbdp')

_trapz = getattr(np, 'trapezoid', None) or np.trapz   # same function; np.trapz (tem_util.py:232) is its deprecated alias

# scipy.special.sph_harm_y(l, 0, ., 0) returns NaN for l >= 646
_SCIPY_LMAX = 645


def zm_latitudes(zm_dlat=1, zm_pole_points=False):
    """tem_diagnostics.py:388-396 — uniform zonal-mean latitude grid (cell midpoints by default)."""
    tol = 1e-6
    assert float(180 / zm_dlat).is_integer(), '180 must be divisible by dlat_out'
    lat_zm = np.arange(-90, 90 + zm_dlat, zm_dlat)
    if lat_zm[-1] > 90 + tol:
        lat_zm = lat_zm[:-1]
    if not zm_pole_points:
        lat_zm = (lat_zm[1:] + lat_zm[:-1]) / 2
    return lat_zm


def sph_basis_recurrence(x, L):
    """Real Y_l^0 = sqrt((2l+1)/4pi) P_l(x) by the normalised three-term recurrence.

    Used only where SciPy cannot evaluate (l >= 646); same recurrence the device kernel uses
    (SURVEY.md §7.2 K1).  x = cos(colatitude).
    """
    x = np.asarray(x, dtype=np.float64)
    Y = np.zeros((x.shape[0], L + 1))
    Y[:, 0] = np.sqrt(1.0 / (4 * np.pi))
    if L >= 1:
        Y[:, 1] = np.sqrt(3.0 / (4 * np.pi)) * x
    for l in range(2, L + 1):
        al = np.sqrt(4.0 * l * l - 1.0) / l
        bl = ((l - 1.0) / l) * np.sqrt((2.0 * l + 1.0) / (2.0 * l - 3.0))
        Y[:, l] = al * x * Y[:, l - 1] - bl * Y[:, l - 2]
    return Y


def sph_basis_dlat(lat_deg, L):
    """d/dphi of the basis of `sph_basis` (phi = latitude in radians), (N, L+1): cos(phi) * d/dx [N_l P_l(x)], x = sin(phi),
    by the differentiated three-term recurrence.  NO reference counterpart: the reference differentiates by
    np.gradient (tem_util.py:154); this restates the optional Legendre-space derivative named by BASELINE.json's
    north_star and is pinned against mpmath in tests/test_oracle_golden.py."""
    x = np.cos(np.deg2rad(90 - np.asarray(lat_deg, dtype=np.float64)))
    c = np.sqrt(np.maximum(0.0, (1.0 - x) * (1.0 + x)))
    n = x.shape[0]
    Y = np.zeros((n, L + 1))
    D = np.zeros((n, L + 1))
    Y[:, 0] = np.sqrt(1.0 / (4.0 * np.pi))
    if L >= 1:
        a1 = np.sqrt(3.0 / (4.0 * np.pi))
        Y[:, 1] = a1 * x
        D[:, 1] = a1
    for l in range(2, L + 1):
        a = np.sqrt(4.0 * l * l - 1.0) / l
        b = ((l - 1.0) / l) * np.sqrt((2.0 * l + 1.0) / (2.0 * l - 3.0))
        Y[:, l] = a * x * Y[:, l - 1] - b * Y[:, l - 2]
        D[:, l] = a * (Y[:, l - 1] + x * D[:, l - 1]) - b * D[:, l - 2]
    return D * c[:, None]


def sph_basis(lat_deg, L):
    """sph_zonal_mean.py:359-363 — Y0[:,l] = sph_harm(0, l, 0, coalt).real, coalt = deg2rad(90-lat).

    `scipy.special.sph_harm(m, n, azimuth, polar)` no longer exists in SciPy >= 1.17; its
    replacement is `sph_harm_y(n, m, polar, azimuth)`.
    """
    lat_deg = np.asarray(lat_deg, dtype=np.float64)
    coalt = np.deg2rad(90 - lat_deg)
    Y0 = np.zeros((lat_deg.shape[0], L + 1))
    for ll in range(0, min(L, _SCIPY_LMAX) + 1):
        Y0[:, ll] = scipy.special.sph_harm_y(ll, 0, coalt, 0).real
    if L > _SCIPY_LMAX:
        Y0[:, _SCIPY_LMAX + 1:] = sph_basis_recurrence(np.cos(coalt), L)[:, _SCIPY_LMAX + 1:]
    return Y0


def sph_matrices(lat, lat_out, L, method='auto', basis='scipy'):
    """sph_zonal_mean.py:358-390 — Y0 (N,L+1), Y0inv (L+1,N), Y0p (M,L+1).

    method='lstsq' is the literal reference call `lstsq(Y0, identity(N))[0]` (needs an N x N
    identity); 'pinv' is the mathematically identical Moore-Penrose pseudo-inverse
    (|lstsq - pinv|_max = 1.7e-17 at N=21,600).  'auto' = lstsq for N <= 6000;
    'normal' = normal equations, for the million-column sampled checks only.
    """
    lat = np.asarray(lat, dtype=np.float64)
    lat_out = np.asarray(lat_out, dtype=np.float64)
    if basis == 'recurrence':
        # O(N L) instead of SciPy's O(N L^2): only for the million-column sampled checks; the recurrence is
        # pinned against SciPy and mpmath in tests/test_oracle_golden.py
        Y0 = sph_basis_recurrence(np.cos(np.deg2rad(90 - lat)), L)
        Y0p = sph_basis_recurrence(np.cos(np.deg2rad(90 - lat_out)), L)
    else:
        Y0 = sph_basis(lat, L)
        Y0p = sph_basis(lat_out, L)
    if method == 'auto':
        method = 'lstsq' if lat.shape[0] <= 6000 else 'pinv'
    if method == 'lstsq':
        Y0inv = scipy.linalg.lstsq(Y0, np.identity(lat.shape[0]))[0]
    elif method == 'pinv':
        Y0inv = np.linalg.pinv(Y0)
    elif method == 'normal':
        # pinv(Y0) = (Y0^T Y0)^-1 Y0^T for full column rank; only for the million-column configs where
        # an SVD of Y0 takes minutes (cond(Y0) <= 23 there)
        Y0inv = scipy.linalg.solve(Y0.T @ Y0, Y0.T, assume_a='pos')
    else:
        raise ValueError(method)
    return Y0, Y0inv, Y0p


def zonal_mean(A, Y, Y0inv, literal=True):
    """sph_zonal_mean.py:243-255 — Abar = (Y @ Y0inv) @ AA with AA = A.reshape(N, DD) (C order).

    literal=False re-associates to Y @ (Y0inv @ AA) (same mathematics, no N x N operator);
    the two agree to <= 1.3e-12 normwise on every TEM output.
    """
    A = np.asarray(A)
    if np.sum(np.isnan(A)) > 0:                                   # sph_zonal_mean.py:219-221
        raise RuntimeError('Variable has nans!')
    shape = A.shape
    N = Y0inv.shape[1]
    if shape[0] != N:                                             # sph_zonal_mean.py:232-237
        raise RuntimeError('Expected the first (leftmost) dimension to be of length %d' % N)
    AA = A.reshape((N, -1)).astype(np.float64, copy=False)
    if literal:
        Abar = np.matmul(np.matmul(Y, Y0inv), AA)
    else:
        Abar = np.matmul(Y, np.matmul(Y0inv, AA))
    return Abar.reshape((Y.shape[0],) + tuple(shape[1:])).astype(A.dtype)


def _ml(A, x):
    """tem_util.py:80 multiply_lat"""
    return np.einsum('ijk,i->ijk', A, x)


def _mp(A, x):
    """tem_util.py:117 multiply_p"""
    return np.einsum('ijk,j->ijk', A, x)


def _p_integral(A, p):
    """tem_util.py:228-232 — out[:,k,:] = trapz(A[:,:k+1,:], p[:k+1], axis=1)"""
    out = np.zeros(A.shape)
    for kk in range(len(p)):
        out[:, kk, :] = _trapz(A[:, :kk + 1, :], p[:kk + 1], axis=1)
    return out


def tem_suite(ua, va, ta, wap, plev_hPa, lat, L=50, zm_dlat=1, p0=None, zm_pole_points=False,
              literal=True, q=None, matrices=None, inv_method='auto'):
    """The whole of `TEMDiagnostics.__init__` plus every diagnostics method.

    Inputs are ndarrays shaped (ncol, plev, time) (the layout the reference transposes to,
    tem_diagnostics.py:342-357); plev in hPa; lat in degrees.  Returns a dict holding every
    intermediate of `_decompose_zm_eddy`/`_compute_fluxes`/`_compute_derivatives`
    (tem_diagnostics.py:510-611) and every method output (tem_diagnostics.py:615-797, tracers
    :801-991), all shaped (M, plev, time) with plev ascending (model top first).
    """
    C = CONSTANTS
    a, Om, H, g0, kap, pi = C['a'], C['Om'], C['H'], C['g0'], C['k'], C['pi']
    if p0 is None:
        p0 = C['P0']
    ua, va, ta, wap = (np.asarray(x) for x in (ua, va, ta, wap))
    dtype = ua.dtype
    plev = np.asarray(plev_hPa, dtype=np.float64)
    qs = [] if q is None else [np.asarray(x) for x in (q if isinstance(q, (list, tuple)) else [q])]
    if ua.ndim == 2:   # reference intends T=1 (tem_diagnostics.py:332-335; broken there)
        ua, va, ta, wap = (x[:, :, None] for x in (ua, va, ta, wap))
        qs = [x[:, :, None] for x in qs]

    # tem_diagnostics.py:372-382 — model top must be the leftmost pressure entry
    if plev[0] > plev[-1]:
        ua, va, ta, wap = (x[:, ::-1, :] for x in (ua, va, ta, wap))
        qs = [x[:, ::-1, :] for x in qs]
        plev = plev[::-1]
    p = plev * 100                                                # :385
    lat_zm = zm_latitudes(zm_dlat, zm_pole_points)                # :388-394
    f = (2 * Om * np.sin(lat_zm * np.pi / 180))[:, np.newaxis, np.newaxis]   # :401,405
    coslat = np.cos(lat_zm * np.pi / 180)                         # :402
    latr = np.deg2rad(lat_zm)

    if matrices is None:
        matrices = sph_matrices(lat, lat_zm, L, method=inv_method)   # :243-248
    Y0, Y0inv, Y0p = matrices
    zm = lambda A: zonal_mean(A, Y0p, Y0inv, literal)             # sph_zonal_mean.py:291-296
    zmn = lambda A: zonal_mean(A, Y0, Y0inv, literal)             # sph_zonal_mean.py:285-290

    out = dict(lat_zm=lat_zm, p=p, plev=plev, coslat=coslat, f=f)
    theta = _mp(ta, (p0 / p) ** kap)                              # :498
    out['theta'] = theta

    # _decompose_zm_eddy :515-529
    ub, up = zm(ua), ua - zmn(ua)
    vb, vp = zm(va), va - zmn(va)
    thetab, thetap = zm(theta), theta - zmn(theta)
    wapb, wapp = zm(wap), wap - zmn(wap)
    # _compute_fluxes :547-557
    upvpb = zm(up * vp)
    upwappb = zm(up * wapp)
    vptpb = zm(vp * thetap)
    # _compute_derivatives :579-599
    dub_dp = np.gradient(ub, p, axis=1)                           # tem_util.py:192
    dthetab_dp = np.gradient(thetab, p, axis=1)
    ubcoslat = _ml(ub, coslat)
    dubcoslat_dlat = np.gradient(ubcoslat, latr, axis=0)          # tem_util.py:154
    psi = vptpb / dthetab_dp                                      # :590
    psicoslat = _ml(psi, coslat)
    dpsicoslat_dlat = np.gradient(psicoslat, latr, axis=0)
    dpsi_dp = np.gradient(psi, p, axis=1)
    int_vbdp = _p_integral(vb, p)                                 # :599
    out.update(ub=ub, vb=vb, thetab=thetab, wapb=wapb, upvpb=upvpb, upwappb=upwappb,
               vptpb=vptpb, dub_dp=dub_dp, dthetab_dp=dthetab_dp, ubcoslat=ubcoslat,
               dubcoslat_dlat=dubcoslat_dlat, psi=psi, psicoslat=psicoslat,
               dpsicoslat_dlat=dpsicoslat_dlat, dpsi_dp=dpsi_dp, int_vbdp=int_vbdp,
               up=up, vp=vp, thetap=thetap, wapp=wapp)

    # diagnostics methods :615-797
    vtem = vb - dpsi_dp                                           # :622
    omegatem = wapb + _ml(dpsicoslat_dlat, 1 / (a * coslat))      # :639
    wtem = _mp(omegatem, -H / p)                                  # :657
    psitem = 2 * pi * a / g0 * _ml(int_vbdp - psi, coslat)        # :674
    x = _ml(dub_dp * psi - upvpb, a * coslat)                     # :691
    epfy = _mp(x, p / p0)                                         # :692
    xz = f - _ml(dubcoslat_dlat, 1 / (a * coslat))                # :709
    epfz = -H / p0 * _ml((xz * psi - upwappb), a * coslat)        # :710
    Fphi = _mp(epfy, p0 / p)                                      # :730
    Fp = epfz * -p0 / H                                           # :731
    Fphicoslat = _ml(Fphi, coslat)                                # :733
    dFphicoslat_dlat = np.gradient(Fphicoslat, latr, axis=0)      # :734
    dFp_dp = np.gradient(Fp, p, axis=1)                           # :735
    epdiv = _ml(dFphicoslat_dlat, 1 / (a * coslat)) + dFp_dp      # :736
    utendepfd = _ml(epdiv, 1 / (a * coslat))                      # :753
    utendvtem = vtem * xz                                         # :771-773
    utendwtem = -omegatem * dub_dp                                # :790-791
    out.update(vtem=vtem, omegatem=omegatem, wtem=wtem, psitem=psitem, epfy=epfy, epfz=epfz,
               epdiv=epdiv, utendepfd=utendepfd, utendvtem=utendvtem, utendwtem=utendwtem)

    # tracers :532-538, :560-570, :602-611, :801-991
    for i, qi in enumerate(qs):
        qb, qp = zm(qi), qi - zmn(qi)
        qpvpb = zm(qp * vp)
        qpwappb = zm(qp * wapp)
        dqb_dp = np.gradient(qb, p, axis=1)
        qbcoslat = _ml(qb, coslat)
        dqbcoslat_dlat = np.gradient(qbcoslat, latr, axis=0)
        etfy = _mp(_ml(dqb_dp * psi - qpvpb, a * coslat), p / p0)                  # :825-826
        xq = -_ml(dqbcoslat_dlat, 1 / (a * coslat))                                # :859
        etfz = -H / p0 * _ml((xq * psi - qpwappb), a * coslat)                     # :860
        Mphi = _mp(etfy, p0 / p)                                                   # :893
        Mp = etfz * -p0 / H                                                        # :894
        dMphicoslat_dlat = np.gradient(_ml(Mphi, coslat), latr, axis=0)            # :896-897
        etdiv = _ml(dMphicoslat_dlat, 1 / (a * coslat)) + np.gradient(Mp, p, axis=1)   # :898-899
        qtendetfd = _ml(etdiv, 1 / (a * coslat))                                   # :928
        qtendvtem = -vtem * _ml(dqbcoslat_dlat, 1 / (a * coslat))                  # :958-959
        qtendwtem = -omegatem * dqb_dp                                             # :986-987
        for nm, v in dict(qb=qb, qpvpb=qpvpb, qpwappb=qpwappb, dqb_dp=dqb_dp, qbcoslat=qbcoslat,
                          dqbcoslat_dlat=dqbcoslat_dlat, etfy=etfy, etfz=etfz, etdiv=etdiv,
                          qtendetfd=qtendetfd, qtendvtem=qtendvtem, qtendwtem=qtendwtem).items():
            out['%s%d' % (nm, i)] = v

    # every public output is cast to the input dtype (:626,643,661,678,696,714,740,757,777,795)
    for nm in list(out):
        if nm not in ('lat_zm', 'p', 'plev', 'coslat', 'f') and isinstance(out[nm], np.ndarray):
            if out[nm].dtype != dtype and nm in TEM_OUTPUTS:
                out[nm] = out[nm].astype(dtype)
    return out
