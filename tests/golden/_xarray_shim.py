"""A minimal stand-in for `xarray` — just enough surface to execute the UNMODIFIED reference source
(/root/reference/PyTEMDiags) in a container where xarray cannot be installed (no network).

Used ONLY by tests/golden/make_golden.py (fixture generation) and by the DataArray-facing API tests as
a DataArray test double.  Semantics implemented: named dims, dimension coordinates, name/attrs,
positional indexing, dim-name-aligned arithmetic, NumPy ufunc/array protocol, and the handful of
methods the reference's hot path calls (SURVEY.md §8c "Oracle B" lists them).
"""
import copy as _copy
import sys
import types

import numpy as np


class DataArray:
    __array_priority__ = 1000

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            coords = dict(data.coords) if coords is None else coords
            name = data.name if name is None else name
            attrs = dict(data.attrs) if attrs is None else attrs
            data = data._v
        self._v = np.asarray(data)
        if dims is None:
            dims = tuple('dim_%d' % i for i in range(self._v.ndim))
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims)
        assert len(self.dims) == self._v.ndim, (self.dims, self._v.shape)
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = np.asarray(v._v if isinstance(v, DataArray) else v)
        self.name = name
        self.attrs = dict(attrs or {})

    # ---- basic protocol ----
    @property
    def values(self):
        return self._v

    @values.setter
    def values(self, v):
        v = np.asarray(v._v if isinstance(v, DataArray) else v)
        assert v.shape == self._v.shape, (v.shape, self._v.shape)
        self._v = v

    shape = property(lambda self: self._v.shape)
    dtype = property(lambda self: self._v.dtype)
    ndim = property(lambda self: self._v.ndim)

    def __len__(self):
        return self._v.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self._v if dtype is None else self._v.astype(dtype)

    def __bool__(self):
        return bool(self._v)

    def __float__(self):
        return float(self._v)

    def __getattr__(self, item):
        attrs = self.__dict__.get('attrs', {})
        if item in attrs:
            return attrs[item]
        raise AttributeError(item)

    def _like(self, v, dims=None, coords=None):
        dims = self.dims if dims is None else dims
        coords = self.coords if coords is None else coords
        coords = {k: c for k, c in coords.items() if k in dims}
        return DataArray(v, coords=coords, dims=dims, name=self.name, attrs=self.attrs)

    def copy(self, deep=True):
        return DataArray(self._v.copy() if deep else self._v, coords=_copy.deepcopy(self.coords), dims=self.dims,
                         name=self.name, attrs=_copy.deepcopy(self.attrs))

    def astype(self, dtype):
        return self._like(self._v.astype(dtype))

    def sum(self, axis=None, **kw):
        return np.sum(self._v, axis=axis)

    def to_netcdf(self, *a, **k):
        return None

    # ---- reshaping ----
    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        return self._like(np.transpose(self._v, order), dims=tuple(dims))

    def expand_dims(self, dim, axis=None):
        axis = self._v.ndim if axis is None else axis
        dims = list(self.dims)
        dims.insert(axis, dim)
        return self._like(np.expand_dims(self._v, axis), dims=tuple(dims))

    def isel(self, **idx):
        v, coords = self._v, dict(self.coords)
        for d, s in idx.items():
            ax = self.dims.index(d)
            sl = [slice(None)] * v.ndim
            sl[ax] = s
            v = v[tuple(sl)]
            if d in coords:
                coords[d] = coords[d][s]
        return self._like(v, coords=coords)

    def rename(self, mapping):
        dims = tuple(mapping.get(d, d) for d in self.dims)
        coords = {mapping.get(k, k): c for k, c in self.coords.items()}
        return self._like(self._v, dims=dims, coords=coords)

    def reindex(self, mapping):
        out = self
        for d, labels in mapping.items():
            labels = np.asarray(labels._v if isinstance(labels, DataArray) else labels)
            pos = {c: i for i, c in enumerate(out.coords[d].tolist())}
            take = [pos[l] for l in labels.tolist()]
            ax = out.dims.index(d)
            coords = dict(out.coords)
            coords[d] = labels.copy()
            out = out._like(np.take(out._v, take, axis=ax), coords=coords)
        return out

    # ---- indexing ----
    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.coords:
                return DataArray(self.coords[key], coords={key: self.coords[key]}, dims=(key,), name=key)
            n = self._v.shape[self.dims.index(key)]
            return DataArray(np.arange(n), dims=(key,), name=key)
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (self._v.ndim - len(key))
        dims, coords = [], {}
        for d, k in zip(self.dims, key):
            if isinstance(k, (int, np.integer)):
                continue
            dims.append(d)
            if d in self.coords:
                coords[d] = self.coords[d][k]
        return DataArray(self._v[key], coords=coords, dims=tuple(dims), name=self.name, attrs=self.attrs)

    def __setitem__(self, key, value):
        if not self._v.flags.writeable:
            self._v = self._v.copy()
        self._v[key] = np.asarray(value._v if isinstance(value, DataArray) else value)

    # ---- arithmetic (dim-name alignment between DataArrays; positional against ndarrays/scalars) ----
    @staticmethod
    def _align(a, b):
        dims = list(a.dims) + [d for d in b.dims if d not in a.dims]

        def put(x):
            v = x._v
            have = [d for d in dims if d in x.dims]
            v = np.transpose(v, [x.dims.index(d) for d in have])
            shape = [v.shape[have.index(d)] if d in have else 1 for d in dims]
            return v.reshape(shape)
        coords = dict(b.coords)
        coords.update(a.coords)
        return put(a), put(b), tuple(dims), coords

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__':
            return NotImplemented
        das = [x for x in inputs if isinstance(x, DataArray)]
        if len(das) == 2 and len(inputs) == 2:
            va, vb, dims, coords = DataArray._align(inputs[0], inputs[1])
            res = ufunc(va, vb, **kwargs)
            return DataArray(res, coords={k: c for k, c in coords.items() if k in dims}, dims=dims,
                             name=das[0].name if das[0].name == das[1].name else None)
        raw = [x._v if isinstance(x, DataArray) else x for x in inputs]
        res = ufunc(*raw, **kwargs)
        ref = das[0]
        if isinstance(res, np.ndarray) and res.ndim == ref.ndim:
            coords = {k: c for k, c in ref.coords.items() if res.shape[ref.dims.index(k)] == len(c)}
            return DataArray(res, coords=coords, dims=ref.dims, name=ref.name, attrs={})
        if isinstance(res, np.ndarray) and res.ndim == 0:
            return DataArray(res, dims=())
        return res

    def _bin(ufunc, reflexive=False):   # noqa: N805
        def op(self, other):
            return ufunc(other, self) if reflexive else ufunc(self, other)
        return op

    __add__ = _bin(np.add); __radd__ = _bin(np.add, True)
    __sub__ = _bin(np.subtract); __rsub__ = _bin(np.subtract, True)
    __mul__ = _bin(np.multiply); __rmul__ = _bin(np.multiply, True)
    __truediv__ = _bin(np.true_divide); __rtruediv__ = _bin(np.true_divide, True)
    __pow__ = _bin(np.power); __rpow__ = _bin(np.power, True)
    __gt__ = _bin(np.greater); __lt__ = _bin(np.less); __ge__ = _bin(np.greater_equal); __le__ = _bin(np.less_equal)

    def __neg__(self):
        return np.negative(self)

    def __repr__(self):
        return '<shim DataArray %s %s %s>' % (self.name, self.dims, self._v.shape)


class _Dataset:
    def to_netcdf(self, *a, **k):
        return None


def open_dataset(path, *a, **k):
    raise FileNotFoundError(path)


def merge(objs, *a, **k):
    return _Dataset()


def install():
    """Register this module as `xarray` (and `xarray.core.dataarray`) in sys.modules."""
    xr = types.ModuleType('xarray')
    core = types.ModuleType('xarray.core')
    da = types.ModuleType('xarray.core.dataarray')
    da.DataArray = DataArray
    core.dataarray = da
    xr.core = core
    xr.DataArray = DataArray
    xr.open_dataset = open_dataset
    xr.merge = merge
    xr.__shim__ = True
    sys.modules['xarray'] = xr
    sys.modules['xarray.core'] = core
    sys.modules['xarray.core.dataarray'] = da
    return xr
