"""Tiny dimensional-analysis stand-in for ``astropy.units`` (test infrastructure).

Only what the reference hot path uses: named units with an SI scale and
(length, mass, time) exponents, ``Quantity`` arithmetic, ``.to()``, ``.value`` and
the ``spectral()`` equivalency (wavelength <-> frequency <-> photon energy).
CODATA-2018 values, matching astropy >= 4.0.
"""
import numpy as np

_H = 6.62607015e-34      # J s
_C = 299792458.0         # m / s


class UnitBase:
    __array_ufunc__ = None   # make ndarray * Unit defer to our __rmul__
    __array_priority__ = 1e6


class Unit(UnitBase):
    def __init__(self, scale, dims):
        self.scale = float(scale)
        self.dims = tuple(dims)          # (m, kg, s, K)

    # unit (*, /, **) unit-or-number
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale,
                        [a + b for a, b in zip(self.dims, other.dims)])
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale / other.scale,
                        [a - b for a, b in zip(self.dims, other.dims)])
        if isinstance(other, Quantity):
            return Quantity(1.0 / other.value, self / other.unit)
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other):
        return Quantity(other, Unit(1.0 / self.scale, [-a for a in self.dims]))

    def __pow__(self, p):
        return Unit(self.scale ** p, [a * p for a in self.dims])

    def is_dimensionless(self):
        return all(abs(a) < 1e-12 for a in self.dims)

    def _factor_to(self, other, equivalencies=None):
        if all(abs(a - b) < 1e-12 for a, b in zip(self.dims, other.dims)):
            return ('lin', self.scale / other.scale)
        if equivalencies == 'spectral':
            return ('spectral', None)
        raise ValueError('incompatible units')


_LEN, _FREQ, _EN = (1, 0, 0, 0), (0, 0, -1, 0), (2, 1, -2, 0)


def _same(d, e):
    return all(abs(a - b) < 1e-12 for a, b in zip(d, e))


class Quantity(UnitBase):
    def __init__(self, value, unit):
        if isinstance(value, Quantity):
            unit = value.unit * unit
            value = value.value
        self.value = value if np.isscalar(value) else np.asarray(value, dtype=float)
        self.unit = unit

    def to(self, unit, equivalencies=None):
        kind, fac = self.unit._factor_to(unit, equivalencies)
        if kind == 'lin':
            return Quantity(self.value * fac, unit)
        # spectral equivalency: go through SI
        v = self.value * self.unit.scale
        src, dst = self.unit.dims, unit.dims
        if _same(src, _LEN):
            hz = _C / v
        elif _same(src, _FREQ):
            hz = v
        elif _same(src, _EN):
            hz = v / _H
        else:
            raise ValueError('not spectral')
        if _same(dst, _LEN):
            out = _C / hz
        elif _same(dst, _FREQ):
            out = hz
        elif _same(dst, _EN):
            out = hz * _H
        else:
            raise ValueError('not spectral')
        return Quantity(out / unit.scale, unit)

    def _coerce(self, other):
        if isinstance(other, Quantity):
            return other
        if isinstance(other, Unit):
            return Quantity(1.0, other)
        return Quantity(other, dimensionless_unscaled)

    def __mul__(self, other):
        o = self._coerce(other)
        return Quantity(self.value * o.value, self.unit * o.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        o = self._coerce(other)
        return Quantity(self.value / o.value, self.unit / o.unit)

    def __rtruediv__(self, other):
        o = self._coerce(other)
        return Quantity(o.value / self.value, o.unit / self.unit)

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def _val_in(self, other):
        o = self._coerce(other)
        return o.to(self.unit).value

    def __add__(self, other):
        return Quantity(self.value + self._val_in(other), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(self.value - self._val_in(other), self.unit)

    def __lt__(self, other):
        return self.value < self._val_in(other)

    def __gt__(self, other):
        return self.value > self._val_in(other)

    def __getitem__(self, k):
        return Quantity(self.value[k], self.unit)

    def __setitem__(self, k, v):
        self.value[k] = self._val_in(v)

    def __len__(self):
        return len(self.value)

    def __float__(self):
        if not self.unit.is_dimensionless():
            raise TypeError('only dimensionless quantities convert to float')
        return float(self.value * self.unit.scale)


def spectral():
    return 'spectral'


dimensionless_unscaled = Unit(1.0, (0, 0, 0, 0))
m = Unit(1.0, (1, 0, 0, 0))
cm = Unit(1e-2, (1, 0, 0, 0))
km = Unit(1e3, (1, 0, 0, 0))
nm = Unit(1e-9, (1, 0, 0, 0))
micron = Unit(1e-6, (1, 0, 0, 0))
angstrom = AA = Angstrom = Unit(1e-10, (1, 0, 0, 0))
au = AU = Unit(1.495978707e11, (1, 0, 0, 0))
pc = Unit(3.0856775814913674e16, (1, 0, 0, 0))
solRad = Rsun = Unit(6.957e8, (1, 0, 0, 0))
jupiterRad = Rjup = R_jup = Unit(7.1492e7, (1, 0, 0, 0))
kg = Unit(1.0, (0, 1, 0, 0))
g = Unit(1e-3, (0, 1, 0, 0))
jupiterMass = Mjup = M_jup = Unit(1.8981245973360505e27, (0, 1, 0, 0))
solMass = Msun = Unit(1.988409870698051e30, (0, 1, 0, 0))
s = Unit(1.0, (0, 0, 1, 0))
Hz = Unit(1.0, (0, 0, -1, 0))
K = Unit(1.0, (0, 0, 0, 1))
J = Unit(1.0, (2, 1, -2, 0))
erg = Unit(1e-7, (2, 1, -2, 0))
eV = Unit(1.602176634e-19, (2, 1, -2, 0))
W = Unit(1.0, (2, 1, -3, 0))
barn = Unit(1e-28, (2, 0, 0, 0))
