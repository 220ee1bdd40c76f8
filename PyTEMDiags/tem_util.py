"""`PyTEMDiags.tem_util` of the alias package: the host-side helpers of the reference's tem_util.py that users call
directly (the stencil helpers - multiply_lat, lat_gradient, p_gradient, p_integral - live in the CUDA epilogue)."""
from pytemdiags_b200.util import format_latlon_arrays, format_latlon_data  # noqa: F401
