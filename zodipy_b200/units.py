"""Minimal unit handling for the host side.

The reference takes ``astropy.units.Quantity`` inputs (``zodipy/model.py:36-45``).  Astropy is an
optional dependency here (it is absent from the build/GPU image): real Astropy quantities are
accepted whenever Astropy is importable, and :class:`Quantity` below is a tiny stand-in with the
handful of attributes the host code needs (``value``, ``unit``, ``isscalar``, ``size``) for the
wavelength / frequency / length units the hot path deals with.
"""
from __future__ import annotations

import numpy as np

C_LIGHT = 299792458.0  # m/s, exact (astropy.constants.c)

_LENGTH = {"m": 1.0, "cm": 1e-2, "mm": 1e-3, "um": 1e-6, "micron": 1e-6, "nm": 1e-9,
           "km": 1e3, "AU": 1.495978707e11, "au": 1.495978707e11}
_FREQ = {"Hz": 1.0, "kHz": 1e3, "MHz": 1e6, "GHz": 1e9, "THz": 1e12}


class UnitConversionError(ValueError):
    """Raised for inconvertible units (mirrors ``astropy.units.UnitConversionError``)."""


def _kind(unit: str) -> str:
    if unit in _LENGTH:
        return "length"
    if unit in _FREQ:
        return "frequency"
    raise UnitConversionError(f"unsupported unit {unit!r}")


class Quantity:
    """Value with a unit string; enough of the Astropy Quantity surface for this package."""

    def __init__(self, value, unit: str):
        _kind(unit)
        self.value = np.asarray(value, dtype=np.float64)
        if self.value.ndim == 0:
            self.value = float(self.value)
        self.unit = unit

    @property
    def isscalar(self) -> bool:
        return np.ndim(self.value) == 0

    @property
    def size(self) -> int:
        return int(np.size(self.value))

    @property
    def ndim(self) -> int:
        return int(np.ndim(self.value))

    @property
    def shape(self):
        return np.shape(self.value)

    def to_value(self, unit: str, spectral: bool = False):
        return convert(self.value, self.unit, unit, spectral)

    def __repr__(self) -> str:
        return f"<Quantity {self.value} {self.unit}>"


def convert(value, src: str, dst: str, spectral: bool = False):
    ks, kd = _kind(src), _kind(dst)
    table_s = _LENGTH if ks == "length" else _FREQ
    table_d = _LENGTH if kd == "length" else _FREQ
    si = np.asarray(value, dtype=np.float64) * table_s[src]
    if ks != kd:
        if not spectral:
            raise UnitConversionError(f"cannot convert {src} to {dst}")
        si = C_LIGHT / si  # units.spectral(): lambda = c / nu
    out = si / table_d[dst]
    return float(out) if np.ndim(out) == 0 else out


def _astropy_units():
    try:
        from astropy import units  # type: ignore
        return units
    except ImportError:
        return None


def is_quantity(x) -> bool:
    if isinstance(x, Quantity):
        return True
    u = _astropy_units()
    return u is not None and isinstance(x, u.Quantity)


def spectral_value(x, unit: str):
    """``x.to_value(unit, equivalencies=spectral())`` for shim or Astropy quantities."""
    if isinstance(x, Quantity):
        return x.to_value(unit, spectral=True)
    u = _astropy_units()
    if u is not None and isinstance(x, u.Quantity):
        try:
            return x.to_value(u.Unit(unit), equivalencies=u.spectral())
        except u.UnitConversionError as err:
            raise UnitConversionError(str(err)) from err
    raise TypeError("expected a Quantity")


def length_value(x, unit: str = "AU"):
    if isinstance(x, Quantity):
        return x.to_value(unit)
    u = _astropy_units()
    if u is not None and isinstance(x, u.Quantity):
        return x.to_value(u.Unit(unit))
    raise TypeError("expected a Quantity")


def native_value(x):
    """The bare number(s) of a quantity in its own unit."""
    return np.asarray(x.value, dtype=np.float64)


def unit_name(x) -> str:
    if isinstance(x, Quantity):
        return x.unit
    return str(x.unit)
