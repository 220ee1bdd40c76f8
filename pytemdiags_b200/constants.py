"""Physical constants of the reference (PyTEMDiags/constants.py:6-14), bit-for-bit.

Set matching DynVarMIP (Gerber & Manzini 2016, section A2).  `pi` is the reference's truncated
3.14159, used by `psitem` (tem_diagnostics.py:674); `np.pi` is used elsewhere, as in the reference.
"""
P0 = 101325      # surface pressure [Pa]
R = 287.058      # gas constant for dry air [J/K/kg]
Cp = 1004.64     # specific heat of dry air at constant pressure [J/K/kg]
g0 = 9.80665     # gravity at mean sea level [m/s^2]
a = 6.37123e6    # radius of Earth [m]
Om = 7.29212e-5  # rotation rate of Earth [1/s]
k = R / Cp       # ratio of gas constant to specific heat
H = 7 * 1e3      # scale height [m]
pi = 3.14159
