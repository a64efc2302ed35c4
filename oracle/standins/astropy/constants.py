"""CODATA-2018 constants as stand-in Quantities (test infrastructure only)."""
from . import units as u

h = u.Quantity(6.62607015e-34, u.J * u.s)
c = u.Quantity(299792458.0, u.m / u.s)
G = u.Quantity(6.6743e-11, u.m ** 3 / u.kg / u.s ** 2)
m_p = u.Quantity(1.67262192369e-27, u.kg)
k_B = u.Quantity(1.380649e-23, u.J / u.K)
