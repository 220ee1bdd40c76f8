"""Synthetic grids and fields for the BASELINE.json configs.

Host (NumPy) generators.  The device generator (`temd_synth_fields`, csrc/temd_fields.cu) uses the
same counter-based hash noise and the same closed forms, so a device-generated slab can be copied
back and fed to the CPU oracle unchanged.  There are no datasets in the container: all benchmark and
parity inputs come from here.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def pg2_grid(ne):
    """Equiangular cubed-sphere physics grid ("pg2"): n = 2*ne cells per face edge, N = 24*ne**2.

    Returns (lat_deg[N], lon_deg[N]).  Face order: 4 equatorial faces, north cap, south cap.
    """
    n = 2 * ne
    ab = -np.pi / 4 + (np.arange(n) + 0.5) * np.pi / (2 * n)
    alpha, beta = np.meshgrid(ab, ab, indexing='ij')
    alpha, beta = alpha.ravel(), beta.ravel()
    lats, lons = [], []
    for face in range(4):
        lats.append(np.arctan(np.tan(beta) * np.cos(alpha)))
        lons.append(alpha + face * np.pi / 2)
    r = np.sqrt(np.tan(alpha) ** 2 + np.tan(beta) ** 2)
    cap = np.arctan2(1.0, r)
    lon_cap = np.arctan2(np.tan(beta), np.tan(alpha))
    lats += [cap, -cap]
    lons += [lon_cap, lon_cap]
    lat = np.concatenate(lats)
    lon = np.mod(np.concatenate(lons), 2 * np.pi)
    return np.rad2deg(lat), np.rad2deg(lon)


def latlon_grid(nlat, nlon, poles=True):
    """Regular lat-lon grid raveled lat-major, ncol = ilat*nlon + ilon (reference
    `tem_util.py:331`, `Dataset.stack(ncol=(lat, lon))`)."""
    if poles:
        lat1 = np.linspace(-90, 90, nlat)
    else:
        d = 180.0 / nlat
        lat1 = -90 + d / 2 + d * np.arange(nlat)
    lon1 = np.arange(nlon) * (360.0 / nlon)
    lat, lon = np.meshgrid(lat1, lon1, indexing='ij')
    return lat.ravel(), lon.ravel()


def default_plev(K):
    """hPa, ascending (model top first)."""
    return np.geomspace(1.0, 1000.0, K)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def hash_noise(seed, field, t, K, N):
    """xi[k, i] in [-sqrt(3), sqrt(3)) (unit variance), keyed on (seed, field, t, k, i)."""
    with np.errstate(over='ignore'):
        key = _splitmix64(np.uint64(seed) * np.uint64(8) + np.uint64(field))
        idx = (np.uint64(t) * np.uint64(K) + np.arange(K, dtype=np.uint64)[:, None]) * np.uint64(N) \
            + np.arange(N, dtype=np.uint64)[None, :]
        z = _splitmix64(idx ^ key)
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (2.0 * u - 1.0) * np.sqrt(3.0)


FIELD_IDS = {'ua': 0, 'va': 1, 'ta': 2, 'wap': 3, 'q': 4}


def synth_fields(lat_deg, lon_deg, plev_hPa, T, seed=0, t0=0, fields=('ua', 'va', 'ta', 'wap'),
                 dtype=np.float64):
    """Statically stable smooth + wave + noise fields, layout [time][lev][ncol] (device layout).

    Returns a dict name -> ndarray (T, K, N).  `t0` offsets the time index so time slabs of one
    long synthetic record can be generated independently.
    """
    phi = np.deg2rad(np.asarray(lat_deg, dtype=np.float64))[None, :]
    lam = np.deg2rad(np.asarray(lon_deg, dtype=np.float64))[None, :]
    s = (np.asarray(plev_hPa, dtype=np.float64) / 1000.0)[:, None]
    K, N = s.shape[0], phi.shape[1]
    c = np.cos(phi)
    out = {f: np.empty((T, K, N), dtype=dtype) for f in fields}
    for it in range(T):
        t = float(t0 + it)
        for f in fields:
            xi = hash_noise(seed, FIELD_IDS[f], t0 + it, K, N)
            if f == 'ta':
                v = 210.0 + 75.0 * s ** 0.19 * c ** 2 + 3.0 * np.sin(3 * lam + 0.3 * t) * c ** 3 * s + 0.5 * xi
            elif f == 'ua':
                v = 30.0 * np.sin(2 * phi) ** 2 * (1.0 - s) + 5.0 * np.cos(4 * lam - 0.2 * t) * c ** 4 + 2.0 * xi
            elif f == 'va':
                v = np.sin(2 * phi) * s + 4.0 * np.sin(4 * lam - 0.2 * t + 0.5) * c ** 4 + 2.0 * xi
            elif f == 'wap':
                v = 0.01 * np.cos(3 * phi) * s + 0.05 * np.sin(3 * lam + 0.3 * t + 1.0) * c ** 3 * s + 0.02 * xi
            elif f == 'q':
                v = 1e-3 * s ** 2 * c ** 2 * (1.0 + 0.3 * np.sin(2 * lam + 0.1 * t)) + 1e-5 * xi
            out[f][it] = v
    return out


# BASELINE.json configs.  grid: ('pg2', ne) or ('latlon', nlat, nlon)
CONFIGS = {
    'config1': dict(grid=('pg2', 30), K=72, T=24, L=50),
    'config2': dict(grid=('pg2', 120), K=72, T=365, L=100),
    'config3': dict(grid=('pg2', 256), K=128, T=96, L=200),
    'config4': dict(grid=('latlon', 721, 1440), K=37, T=240, L=300),
    'config5': dict(grid=('pg2', 120), K=72, T=24, L=(25, 50, 100, 200, 400, 800)),
}


def make_grid(spec):
    if spec[0] == 'pg2':
        return pg2_grid(spec[1])
    if spec[0] == 'latlon':
        return latlon_grid(spec[1], spec[2])
    raise ValueError(spec)
