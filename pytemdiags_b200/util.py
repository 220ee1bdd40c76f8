"""Host-side helpers mirroring the non-hot-path utilities of reference PyTEMDiags/tem_util.py."""
import numpy as np


def format_latlon_data(A, lat, lon, lat_axis=-2, lon_axis=-1):
    """Ravel structured (..., lat, lon) data into the unstructured `ncol` layout the averager expects.

    NumPy analogue of `tem_util.format_latlon_data` (tem_util.py:247-342), which does
    `Dataset.stack(ncol=(lat, lon))`: the column order is latitude-major, `ncol = ilat * NLON + ilon` (:331).

    Returns (A_ncol, lat_ncol, lon_ncol): `A_ncol` has the lat and lon axes replaced by one trailing `ncol` axis,
    `lat_ncol` / `lon_ncol` give the coordinates of every column (degrees, length NLAT*NLON).
    """
    A = np.asarray(A)
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    la, lo = lat_axis % A.ndim, lon_axis % A.ndim
    if la == lo or A.shape[la] != lat.shape[0] or A.shape[lo] != lon.shape[0]:
        raise RuntimeError('lat / lon axes of the data do not match the coordinate vectors')
    rest = [ax for ax in range(A.ndim) if ax not in (la, lo)]
    A2 = np.transpose(A, rest + [la, lo])
    A2 = np.ascontiguousarray(A2).reshape(A2.shape[:-2] + (lat.shape[0] * lon.shape[0],))
    lat_ncol = np.repeat(lat, lon.shape[0])
    lon_ncol = np.tile(lon, lat.shape[0])
    return A2, lat_ncol, lon_ncol
