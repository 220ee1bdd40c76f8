"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference source
(/root/reference/PyTEMDiags) in the build container.

The reference cannot be imported as-is here: `xarray` is not installed (no network) and
`scipy.special.sph_harm` was removed from the installed SciPy.  This script injects (a) the xarray
stand-in of tests/golden/_xarray_shim.py and (b) `sph_harm(m, n, az, polar) := sph_harm_y(n, m, polar, az)`
into the import system, then imports the reference package from /root/reference and calls its public
API (`PyTEMDiags.TEMDiagnostics`, `PyTEMDiags.sph_zonal_averager`) on small synthetic inputs.
No reference source is copied into this repo; /root/reference is only read.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

METHODS = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')
PROPS = ('ub', 'vb', 'thetab', 'wapb', 'upvpb', 'upwappb', 'vptpb', 'dub_dp', 'dthetab_dp', 'ubcoslat',
         'dubcoslat_dlat', 'psi', 'psicoslat', 'dpsicoslat_dlat', 'dpsi_dp', 'int_vbdp')
TRACER_METHODS = ('etfy', 'etfz', 'etdiv', 'qtendetfd', 'qtendvtem', 'qtendwtem')
TRACER_PROPS = ('qb', 'qpvpb', 'qpwappb', 'dqb_dp', 'qbcoslat', 'dqbcoslat_dlat')


def import_reference():
    import _xarray_shim
    xr = _xarray_shim.install()
    import scipy.special as sp
    if not hasattr(sp, 'sph_harm'):
        sp.sph_harm = lambda m, n, theta, phi: sp.sph_harm_y(n, m, phi, theta)
    sys.path.insert(0, '/root/reference')
    import PyTEMDiags
    assert PyTEMDiags.__file__.startswith('/root/reference'), PyTEMDiags.__file__
    return PyTEMDiags, xr


CASES = {
    # name: grid, K, T, L, kwargs, dims order handed to the reference, plev order, tracer?
    'tem_pg2_ne5_L16': dict(grid=('pg2', 5), K=10, T=3, L=16, seed=11, dims=('ncol', 'plev', 'time'), flip=False, q=False, kw={}),
    'tem_pg2_ne4_L12_flipped_tracer': dict(grid=('pg2', 4), K=8, T=2, L=12, seed=12, dims=('time', 'plev', 'ncol'), flip=True, q=True, kw={}),
    'tem_latlon_24x48_L14_dlat2': dict(grid=('latlon', 24, 48), K=6, T=2, L=14, seed=13, dims=('plev', 'ncol', 'time'), flip=False, q=False,
                                       kw=dict(zm_dlat=2)),
}


def make_case(name, spec, PyTEMDiags, xr):
    from pytemdiags_b200 import synthetic as syn
    lat, lon = syn.make_grid(spec['grid']) if spec['grid'][0] == 'pg2' else syn.latlon_grid(spec['grid'][1], spec['grid'][2], poles=False)
    K, T, L = spec['K'], spec['T'], spec['L']
    plev = syn.default_plev(K)
    fields = ('ua', 'va', 'ta', 'wap') + (('q',) if spec['q'] else ())
    f = syn.synth_fields(lat, lon, plev, T, seed=spec['seed'], fields=fields)      # [T][K][N], plev ascending
    plev_in = plev[::-1].copy() if spec['flip'] else plev
    time = np.arange(T) * 6.0
    das = {}
    for n in fields:
        a = f[n][:, ::-1, :] if spec['flip'] else f[n]
        a = np.ascontiguousarray(np.transpose(a, [('time', 'plev', 'ncol').index(d) for d in spec['dims']]))
        das[n] = xr.DataArray(a, dims=spec['dims'], coords={'plev': plev_in, 'time': time}, name=n)
    lat_da = xr.DataArray(lat, dims=('ncol',), name='lat')
    tem = PyTEMDiags.TEMDiagnostics(das['ua'], das['va'], das['ta'], das['wap'], lat_da,
                                    q=das.get('q'), L=L, debug_level=0, **spec['kw'])
    out = dict(lat=lat, lon=lon, plev_in=plev_in, time=time, L=np.int64(L), dims=np.array(spec['dims']),
               lat_zm=np.asarray(tem.lat), p=np.asarray(tem.p.values), zm_dlat=np.float64(spec['kw'].get('zm_dlat', 1)))
    for n in fields:
        out['in_' + n] = np.asarray(das[n].values)
    for m in METHODS:
        r = getattr(tem, m)()
        assert r.dims == ('lat', 'plev', 'time'), (m, r.dims)
        out['ref_' + m] = np.asarray(r.values)
    for p_ in PROPS:
        out['ref_' + p_] = np.asarray(getattr(tem, p_).values)
    if spec['q']:
        for m in TRACER_METHODS:
            out['ref_' + m + '0'] = np.asarray(getattr(tem, m)(0).values)
        for p_ in TRACER_PROPS:
            out['ref_' + p_ + '0'] = np.asarray(getattr(tem, p_)[0].values)
    # the averager's matrices and one stand-alone zonal mean (sph_zonal_mean.py:285-296)
    out['ref_Y0inv'] = np.asarray(tem.ZM.Y0inv)
    A = das['ua'].transpose('ncol', 'plev', 'time')
    out['ref_zm_ua'] = np.asarray(tem.ZM.sph_zonal_mean(A).values)
    out['ref_zmnative_ua'] = np.asarray(tem.ZM.sph_zonal_mean_native(A).values)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.2f MB' % (os.path.getsize(path) / 1e6))


def main():
    PyTEMDiags, xr = import_reference()
    for name, spec in CASES.items():
        make_case(name, spec, PyTEMDiags, xr)


if __name__ == '__main__':
    main()
