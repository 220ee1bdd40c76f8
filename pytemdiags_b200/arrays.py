"""Input/output adapters: xarray-style DataArrays (duck-typed), numpy arrays and torch tensors.

The reference accepts only `xarray.DataArray` (tem_diagnostics.py:311-313, sph_zonal_mean.py:217-218).
xarray is not a dependency here: anything exposing `.dims`, `.values` (and optionally `.coords`,
`.name`, `.attrs`) is treated as a DataArray and outputs are rebuilt with the same class; raw
numpy / torch inputs return raw numpy / torch outputs.  This is host plumbing only — no arithmetic
on field data happens here beyond layout changes and dtype casts.
"""
import numpy as np
import torch


def is_dataarray(x):
    return hasattr(x, 'dims') and hasattr(x, 'values') and not isinstance(x, (np.ndarray, torch.Tensor))


def kind_of(x):
    if is_dataarray(x):
        return 'dataarray'
    if isinstance(x, torch.Tensor):
        return 'torch'
    return 'numpy'


def raw(x):
    """The underlying ndarray / tensor of an input."""
    if is_dataarray(x):
        v = x.values
        return v if isinstance(v, (np.ndarray, torch.Tensor)) else np.asarray(v)
    if isinstance(x, torch.Tensor):
        return x
    return np.asarray(x)


def coord_values(x, dim):
    """1-D coordinate values of a DataArray-like along `dim`, or None."""
    if not is_dataarray(x):
        return None
    try:
        c = x.coords[dim]
    except Exception:
        try:
            c = x[dim]
        except Exception:
            return None
    v = getattr(c, 'values', c)
    return np.asarray(v)


def dtype_of(x):
    r = raw(x)
    if isinstance(r, torch.Tensor):
        return {torch.float64: np.dtype('float64'), torch.float32: np.dtype('float32'),
                torch.float16: np.dtype('float16')}.get(r.dtype, np.dtype('float64'))
    return r.dtype


def to_device_f64(r, device, non_blocking=False):
    """ndarray / tensor (any float dtype, any device) -> float64 tensor on `device` (same shape/strides order)."""
    if isinstance(r, np.ndarray):
        if r.dtype.byteorder not in ('=', '|') or not r.flags.writeable:
            r = np.array(r, dtype=r.dtype.newbyteorder('='))
        r = torch.from_numpy(r)
    t = r.to(device, non_blocking=non_blocking)
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t


def from_device(t, like_kind, dtype, device_of_input=None):
    """Device float64 tensor -> output container matching the input kind / dtype."""
    if like_kind == 'torch':
        tdt = {np.dtype('float64'): torch.float64, np.dtype('float32'): torch.float32,
               np.dtype('float16'): torch.float16}.get(np.dtype(dtype), torch.float64)
        out = t.to(tdt)
        if device_of_input is not None and out.device != device_of_input:
            out = out.to(device_of_input)
        return out
    return t.cpu().numpy().astype(dtype, copy=False)


def make_dataarray(like, data, dims, coords=None, name=None, attrs=None):
    """Rebuild a DataArray of `like`'s class (xarray.DataArray signature: data, coords, dims, name, attrs)."""
    cls = like.__class__
    return cls(data, coords=coords, dims=dims, name=name, attrs=dict(attrs or {}))
