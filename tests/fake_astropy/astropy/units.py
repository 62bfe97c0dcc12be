import numpy as np

C_LIGHT = 299792458.0


class UnitConversionError(ValueError):
    pass


class UnitBase:
    __array_ufunc__ = None  # make ndarray operators defer to __rmul__ / __rlshift__ (as Astropy's units do)

    def __init__(self, name, kind, scale):
        self.name, self.kind, self.scale = name, kind, scale

    def __str__(self):
        return self.name

    __repr__ = __str__

    def __truediv__(self, other):
        return UnitBase(f"{self.name} / {other.name}", f"{self.kind}/{other.kind}", self.scale / other.scale)

    def __rmul__(self, value):
        return Quantity(value, self)

    def __rlshift__(self, value):
        return Quantity(value, self)

    def __eq__(self, other):
        return isinstance(other, UnitBase) and (self.kind, self.scale) == (other.kind, other.scale)

    def __hash__(self):
        return hash((self.kind, self.scale))


_REGISTRY = {}


def _u(name, kind, scale, *aliases):
    unit = UnitBase(name, kind, scale)
    for n in (name, *aliases):
        _REGISTRY[n] = unit
    return unit


m = _u("m", "length", 1.0)
cm = _u("cm", "length", 1e-2)
km = _u("km", "length", 1e3)
micron = um = _u("micron", "length", 1e-6, "um")
AU = _u("AU", "length", 1.495978707e11, "au")
Hz = _u("Hz", "frequency", 1.0)
GHz = _u("GHz", "frequency", 1e9)
deg = _u("deg", "angle", np.pi / 180)
rad = _u("rad", "angle", 1.0)
s = _u("s", "time", 1.0)
hour = _u("hour", "time", 3600.0)
day = _u("day", "time", 86400.0)
K = _u("K", "temperature", 1.0)
MJy = _u("MJy", "sfd", 1e6)
Jy = _u("Jy", "sfd", 1.0)
sr = _u("sr", "solid_angle", 1.0)


def Unit(name):
    if isinstance(name, UnitBase):
        return name
    try:
        return _REGISTRY[str(name)]
    except KeyError as err:
        raise ValueError(f"unknown unit {name!r}") from err


def spectral():
    return "spectral"


class Quantity(np.ndarray):
    def __new__(cls, value, unit=None):
        obj = np.asarray(value, dtype=np.float64).view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", None)

    @property
    def value(self):
        return np.asarray(self).copy() if self.ndim else float(np.asarray(self))

    @property
    def isscalar(self):
        return self.ndim == 0

    def to(self, unit, equivalencies=None):
        unit = Unit(unit)
        if unit.kind == self.unit.kind:
            return Quantity(np.asarray(self) * (self.unit.scale / unit.scale), unit)
        if {unit.kind, self.unit.kind} == {"length", "frequency"} and equivalencies == "spectral":
            return Quantity((C_LIGHT / (np.asarray(self) * self.unit.scale)) / unit.scale, unit)
        raise UnitConversionError(f"'{self.unit}' and '{unit}' are not convertible")

    def to_value(self, unit, equivalencies=None):
        return self.to(unit, equivalencies).value

    def __lshift__(self, unit):
        return Quantity(np.asarray(self), unit)

    def sum(self, *a, **k):
        return Quantity(np.asarray(self).sum(*a, **k), self.unit)
