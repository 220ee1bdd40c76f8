"""Alias package so `import PyTEMDiags` resolves to the B200-native implementation (drop-in for
jhollowed/PyTEMDiags: same public names as the reference's PyTEMDiags/__init__.py:12-16)."""
from pytemdiags_b200 import TEMDiagnostics, sph_zonal_averager  # noqa: F401
from . import tem_util  # noqa: F401

__version__ = '0.1'
