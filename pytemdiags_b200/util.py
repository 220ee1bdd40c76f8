"""Host-side helpers mirroring the non-hot-path utilities of reference PyTEMDiags/tem_util.py.

`format_latlon_data` keeps the reference's signature (tem_util.py:247-249): Dataset in, Dataset out.  A "Dataset" here
is either a real `xarray.Dataset` (handled with xarray's own stack / drop_vars, exactly the reference's statements)
or, where xarray is not installed, a plain mapping  name -> (dims, ndarray)  which gets the same treatment in NumPy.
`format_latlon_arrays` is the explicit array-level variant.  Pure data movement: no arithmetic on field data.
"""
import numpy as np


def format_latlon_arrays(A, lat, lon, lat_axis=-2, lon_axis=-1):
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


def _midpoint_bounds(x):
    """Cell bounds at the midpoints between neighbours (tem_util.py:311-312, 322-323)."""
    x = np.asarray(x, dtype=np.float64)
    d = np.diff(np.hstack([x, x[-1] + (x[-1] - x[-2])]))
    return np.vstack([x - d / 2, x + d / 2]).T


def format_latlon_data(data, *args, lat_name='lat', lon_name='lon', latbnd_name='lat_bnds', lonbnd_name='lon_bnds',
                       bnddim_name='nbnd', **kw):
    """Reference signature (tem_util.py:247-249): `format_latlon_data(data, lat_name='lat', lon_name='lon',
    latbnd_name='lat_bnds', lonbnd_name='lon_bnds', bnddim_name='nbnd')` -> Dataset whose (lat, lon) dimensions
    are stacked into one leading `ncol` dimension (lat-major, :331) with `lat` / `lon` kept as `('ncol',)` variables
    and cell-bound variables added when missing (:309-328).

    `data`: an `xarray.Dataset`, or a mapping  name -> (dims, array)  (returned as a new dict of the same form).
    Called with bare arrays - `format_latlon_data(A, lat, lon, lat_axis=-2, lon_axis=-1)` - it forwards to
    `format_latlon_arrays` (the round-1 calling convention of this build)."""
    if isinstance(data, np.ndarray) or (args and not isinstance(args[0], str)):
        return format_latlon_arrays(data, *args, **kw)
    if kw:
        raise TypeError('unexpected keyword arguments: {}'.format(sorted(kw)))
    names = dict(zip(('lat_name', 'lon_name', 'latbnd_name', 'lonbnd_name', 'bnddim_name'), args))
    lat_name, lon_name = names.get('lat_name', lat_name), names.get('lon_name', lon_name)
    latbnd_name, lonbnd_name = names.get('latbnd_name', latbnd_name), names.get('lonbnd_name', lonbnd_name)
    bnddim_name = names.get('bnddim_name', bnddim_name)

    if hasattr(data, 'stack') and hasattr(data, 'drop_vars'):
        # a real xarray.Dataset: the reference's own statements (tem_util.py:304-342)
        lat, lon = data[lat_name], data[lon_name]
        for bname, coord in ((latbnd_name, lat), (lonbnd_name, lon)):
            if bname not in data.variables:
                data[bname] = (coord.dims + (bnddim_name,), _midpoint_bounds(coord.values))
            elif bnddim_name not in data[bname].dims:
                raise RuntimeError('Variable {} does not have dimension {}. Dimensions are: {}. Did you specify the '
                                   'latbnd_name, lonbnd_name, and bnddim_name arguments to format_latlon_data() '
                                   'correctly?'.format(bname, bnddim_name, data[bname].dims))
        data = data.stack(ncol=(lat_name, lon_name)).transpose('ncol', ...)
        lats, lons = data[lat_name].values, data[lon_name].values
        data = data.drop_vars((lat_name, lon_name))
        data[lat_name] = ('ncol', lats)
        data[lon_name] = ('ncol', lons)
        return data

    # mapping name -> (dims, array)
    src = {k: (tuple(v[0]), np.asarray(v[1])) for k, v in dict(data).items()}
    if lat_name not in src or lon_name not in src:
        raise RuntimeError('format_latlon_data: the dataset needs coordinate variables {} and {}'.format(lat_name, lon_name))
    lat, lon = np.asarray(src[lat_name][1], dtype=np.float64), np.asarray(src[lon_name][1], dtype=np.float64)
    for bname, cname, coord in ((latbnd_name, lat_name, lat), (lonbnd_name, lon_name, lon)):
        if bname not in src:
            src[bname] = ((cname, bnddim_name), _midpoint_bounds(coord))
        elif bnddim_name not in src[bname][0]:
            raise RuntimeError('Variable {} does not have dimension {}. Dimensions are: {}. Did you specify the '
                               'latbnd_name, lonbnd_name, and bnddim_name arguments to format_latlon_data() '
                               'correctly?'.format(bname, bnddim_name, src[bname][0]))
    out = {}
    nlat, nlon = lat.shape[0], lon.shape[0]
    for name, (dims, arr) in src.items():
        if name in (lat_name, lon_name):
            continue
        has_lat, has_lon = lat_name in dims, lon_name in dims
        if not (has_lat or has_lon):
            out[name] = (dims, arr)
            continue
        # xarray's stack broadcasts a variable that carries only one of the two dimensions
        if not has_lat:
            arr, dims = np.broadcast_to(arr[..., None], arr.shape + (nlat,)), dims + (lat_name,)
        if not has_lon:
            arr, dims = np.broadcast_to(arr[..., None], arr.shape + (nlon,)), dims + (lon_name,)
        a2, _, _ = format_latlon_arrays(arr, lat, lon, dims.index(lat_name), dims.index(lon_name))
        rest = tuple(d for d in dims if d not in (lat_name, lon_name))
        out[name] = (('ncol',) + rest, np.ascontiguousarray(np.moveaxis(a2, -1, 0)))      # .transpose('ncol', ...)
    out[lat_name] = (('ncol',), np.repeat(lat, nlon))
    out[lon_name] = (('ncol',), np.tile(lon, nlat))
    return out
