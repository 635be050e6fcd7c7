"""
Minimal physical-units layer with the subset of the ``simtk.unit`` interface that the
atomsmm API passes across its boundary (reference call sites: forces.py:407,446,497;
propagators.py:717,1188-1191,1253-1254; computers.py:71,88,123; utils.py:16).

Internally every unit is a scale factor relative to SI(+mol) and a vector of dimension
exponents (m, s, kg, K, mol, C, rad).  The "MD unit system" used by the engine is
nm / ps / dalton (= g/mol) / K / e / rad, in which kJ/mol == dalton*nm^2/ps^2 == 1.

Physical constants are CODATA-2006, the values simtk.unit carries: they are
what the reference's pressure goldens contain.
"""

import math
from fractions import Fraction

import numpy as np

_NDIM = 7  # m, s, kg, K, mol, C, rad
_E_CHARGE = 1.602176487e-19
_MD_BASE = (1e-9, 1e-12, 1e-3, 1.0, 1.0, _E_CHARGE, 1.0)


def _dims(**kw):
    order = ('m', 's', 'kg', 'K', 'mol', 'C', 'rad')
    return tuple(Fraction(kw.get(k, 0)) for k in order)


_ZERO = _dims()


class Unit(object):
    __slots__ = ('factor', 'dims', 'name')
    __array_priority__ = 100

    def __init__(self, factor, dims=_ZERO, name=None):
        self.factor = float(factor)
        self.dims = tuple(dims)
        self.name = name

    # -- algebra ---------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.factor*other.factor, tuple(a + b for a, b in zip(self.dims, other.dims)))
        if isinstance(other, Quantity):
            return Quantity(other._value, self*other.unit)._reduce()
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.factor/other.factor, tuple(a - b for a, b in zip(self.dims, other.dims)))
        if isinstance(other, Quantity):
            return Quantity(1.0/other._value, self/other.unit)._reduce()
        return Quantity(1.0/other, self)

    def __rtruediv__(self, other):
        return Quantity(other, Unit(1.0/self.factor, tuple(-a for a in self.dims)))

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(12)
        return Unit(self.factor**float(p), tuple(a*p for a in self.dims))

    def sqrt(self):
        return self**Fraction(1, 2)

    def is_dimensionless(self):
        return all(d == 0 for d in self.dims)

    def is_compatible(self, other):
        return self.dims == other.dims

    def md_factor(self):
        """Multiply a value expressed in this unit by this to get MD-system units."""
        f = self.factor
        for d, base in zip(self.dims, _MD_BASE):
            if d != 0:
                f /= base**float(d)
        return f

    def conversion_factor_to(self, other):
        if self.dims != other.dims:
            raise TypeError('incompatible units: %r vs %r' % (self, other))
        return self.factor/other.factor

    def __eq__(self, other):
        return isinstance(other, Unit) and self.dims == other.dims and \
            math.isclose(self.factor, other.factor, rel_tol=1e-12)

    def __hash__(self):
        return hash(self.dims)

    def __repr__(self):
        if self.name:
            return self.name
        names = ('m', 's', 'kg', 'K', 'mol', 'C', 'rad')
        body = '*'.join('%s^%s' % (n, d) for n, d in zip(names, self.dims) if d != 0)
        return 'Unit(%g %s)' % (self.factor, body or '1')


class Quantity(object):
    __slots__ = ('_value', 'unit')
    __array_priority__ = 99

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity):
            unit = value.unit if unit is None else unit
            value = value.value_in_unit(unit)
        elif isinstance(value, (list, tuple)) and len(value) > 0 and isinstance(value[0], Quantity):
            unit = value[0].unit if unit is None else unit
            value = [v.value_in_unit(unit) for v in value]
        self._value = value
        self.unit = dimensionless if unit is None else unit

    def _reduce(self):
        if self.unit.is_dimensionless():
            return self._value*self.unit.factor
        return self

    def value_in_unit(self, unit):
        f = self.unit.conversion_factor_to(unit)
        v = self._value
        if isinstance(v, (list, tuple)):
            v = np.asarray(v, dtype=float)
        return v if f == 1.0 else v*f

    def in_units_of(self, unit):
        return Quantity(self.value_in_unit(unit), unit)

    def value_in_md_units(self):
        v = self._value
        if isinstance(v, (list, tuple)):
            v = np.asarray(v, dtype=float)
        return v*self.unit.md_factor()

    # -- algebra ---------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self._value*other._value, self.unit*other.unit)._reduce()
        if isinstance(other, Unit):
            return Quantity(self._value, self.unit*other)._reduce()
        return Quantity(self._value*other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self._value/other._value, self.unit/other.unit)._reduce()
        if isinstance(other, Unit):
            return Quantity(self._value, self.unit/other)._reduce()
        return Quantity(self._value/other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other/self._value, Unit(1.0)/self.unit)

    def __pow__(self, p):
        return Quantity(self._value**p, self.unit**p)._reduce()

    def sqrt(self):
        return Quantity(np.sqrt(self._value) if isinstance(self._value, np.ndarray) else math.sqrt(self._value),
                        self.unit.sqrt())._reduce()

    def __add__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self._value + other.value_in_unit(self.unit), self.unit)
        if other == 0:
            return self
        raise TypeError('cannot add a Quantity and a bare number')

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self._value - other.value_in_unit(self.unit), self.unit)
        raise TypeError('cannot subtract a bare number from a Quantity')

    def __rsub__(self, other):
        return (-self).__add__(other)

    def __neg__(self):
        return Quantity(-self._value, self.unit)

    def __abs__(self):
        return Quantity(abs(self._value), self.unit)

    def _cmp(self, other):
        if isinstance(other, Quantity):
            return self._value, other.value_in_unit(self.unit)
        return self._value*self.unit.md_factor(), other

    def __lt__(self, other):
        a, b = self._cmp(other)
        return a < b

    def __le__(self, other):
        a, b = self._cmp(other)
        return a <= b

    def __gt__(self, other):
        a, b = self._cmp(other)
        return a > b

    def __ge__(self, other):
        a, b = self._cmp(other)
        return a >= b

    def __eq__(self, other):
        if isinstance(other, Quantity):
            if self.unit.dims != other.unit.dims:
                return False
            a, b = self._cmp(other)
            return bool(np.all(a == b))
        return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((float(self._value) if np.isscalar(self._value) else id(self), self.unit.dims))

    def __len__(self):
        return len(self._value)

    def __getitem__(self, i):
        return Quantity(self._value[i], self.unit)

    def __iter__(self):
        for v in self._value:
            yield Quantity(v, self.unit)

    def __float__(self):
        return float(self._value*self.unit.md_factor())

    def __repr__(self):
        return 'Quantity(value=%r, unit=%r)' % (self._value, self.unit)

    def __str__(self):
        return '%s %r' % (self._value, self.unit)


def is_quantity(x):
    return isinstance(x, Quantity)


def sqrt(x):
    if isinstance(x, Quantity):
        return x.sqrt()
    return math.sqrt(x)


def md_value(x):
    """Strip units: Quantity -> number/array in MD units (nm, ps, dalton, kJ/mol, K, e, rad)."""
    if isinstance(x, Quantity):
        return x.value_in_md_units()
    if isinstance(x, Unit):
        return x.md_factor()
    return x


dimensionless = Unit(1.0, _ZERO, 'dimensionless')

meter = meters = Unit(1.0, _dims(m=1), 'meter')
nanometer = nanometers = Unit(1e-9, _dims(m=1), 'nanometer')
angstrom = angstroms = Unit(1e-10, _dims(m=1), 'angstrom')
second = seconds = Unit(1.0, _dims(s=1), 'second')
picosecond = picoseconds = Unit(1e-12, _dims(s=1), 'picosecond')
femtosecond = femtoseconds = Unit(1e-15, _dims(s=1), 'femtosecond')
nanosecond = nanoseconds = Unit(1e-9, _dims(s=1), 'nanosecond')
day = days = Unit(86400.0, _dims(s=1), 'day')
kilogram = kilograms = Unit(1.0, _dims(kg=1), 'kilogram')
gram = grams = Unit(1e-3, _dims(kg=1), 'gram')
mole = moles = Unit(1.0, _dims(mol=1), 'mole')
dalton = daltons = amu = amus = Unit(1e-3, _dims(kg=1, mol=-1), 'dalton')
kelvin = kelvins = Unit(1.0, _dims(K=1), 'kelvin')
coulomb = coulombs = Unit(1.0, _dims(C=1), 'coulomb')
elementary_charge = elementary_charges = Unit(_E_CHARGE, _dims(C=1), 'elementary charge')
radian = radians = Unit(1.0, _dims(rad=1), 'radian')
degree = degrees = Unit(math.pi/180.0, _dims(rad=1), 'degree')
joule = joules = Unit(1.0, _dims(kg=1, m=2, s=-2), 'joule')
kilojoule = kilojoules = Unit(1e3, _dims(kg=1, m=2, s=-2), 'kilojoule')
kilojoule_per_mole = kilojoules_per_mole = Unit(1e3, _dims(kg=1, m=2, s=-2, mol=-1), 'kilojoule/mole')
kilocalorie_per_mole = kilocalories_per_mole = Unit(4184.0, _dims(kg=1, m=2, s=-2, mol=-1), 'kilocalorie/mole')
pascal = pascals = Unit(1.0, _dims(kg=1, m=-1, s=-2), 'pascal')
bar = bars = Unit(1e5, _dims(kg=1, m=-1, s=-2), 'bar')
atmosphere = atmospheres = Unit(101325.0, _dims(kg=1, m=-1, s=-2), 'atmosphere')
item = items = Unit(1.0, _ZERO, 'item')

BOLTZMANN_CONSTANT_kB = Quantity(1.3806504e-23, joule/kelvin)
AVOGADRO_CONSTANT_NA = Quantity(6.02214179e23, Unit(1.0)/mole)
MOLAR_GAS_CONSTANT_R = BOLTZMANN_CONSTANT_kB*AVOGADRO_CONSTANT_NA
