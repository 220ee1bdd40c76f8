"""pytemdiags_b200 — B200-native zonal-mean + Transformed Eulerian Mean pipeline.

Drop-in for the hot path of jhollowed/PyTEMDiags: `TEMDiagnostics` and `sph_zonal_averager` keep the
reference's constructor arguments, methods and error behaviour (PyTEMDiags/__init__.py:12-13); all
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of `include/temd.h`.
"""
from .zonal import sph_zonal_averager  # noqa: F401
from .tem import TEMDiagnostics, release_host_staging  # noqa: F401

__version__ = '0.1'
