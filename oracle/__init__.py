"""CPU oracle for the PyTEMDiags zonal-mean + TEM hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 NumPy/SciPy restatement of the
reference algorithm (jhollowed/PyTEMDiags, `PyTEMDiags/sph_zonal_mean.py`,
`PyTEMDiags/tem_diagnostics.py`, `PyTEMDiags/tem_util.py`, `PyTEMDiags/constants.py`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  The product package `pytemdiags_b200` never imports it and has no CPU
fallback.

Parity pinning: the reference ships no golden vectors.  The oracle is pinned
against outputs of the UNMODIFIED reference source executed in the build container through an
`xarray` shim (`tests/golden/make_golden.py`, fixtures `tests/golden/*.npz`) and against the
analytic known-answer checks of the reference's `tests_sph_zonal_mean.py:297-477`.
"""
from .tem_oracle import (  # noqa: F401
    CONSTANTS, zm_latitudes, sph_basis, sph_basis_dlat, sph_basis_recurrence, sph_matrices,
    zonal_mean, tem_suite, TEM_OUTPUTS, TEM_INTERMEDIATES,
)
