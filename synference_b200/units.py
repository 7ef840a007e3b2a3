"""Minimal unit handling for the mock-library hot path.

The reference API hands around ``unyt`` quantities (``library.py:46``;
``Myr``, ``Msun``, ``Jy`` ...).  ``unyt`` is not installable in this image, so
this module provides the small subset the hot path needs with the same
spelling: ``value * Unit`` builds a :class:`Quantity`, ``Quantity.to(unit)``,
``.to_value(unit)``, ``.value`` and ``.units``.  Objects coming from a real
``unyt`` install are accepted by duck typing (``.value`` / ``str(.units)``).
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "Unit", "Quantity", "unyt_array", "unyt_quantity", "as_quantity", "strip_units",
    "yr", "Myr", "Gyr", "Msun", "Jy", "mJy", "uJy", "nJy", "Angstrom", "um", "nm",
    "Mpc", "cm", "dimensionless",
]

# name -> (dimension, factor to the base unit of that dimension)
_REGISTRY = {
    "yr": ("time", 1.0), "Myr": ("time", 1.0e6), "Gyr": ("time", 1.0e9),
    "Msun": ("mass", 1.0), "Msun/yr": ("mass_rate", 1.0), "K": ("temperature", 1.0),
    "erg/s/Hz": ("luminosity_density", 1.0),
    "Jy": ("flux", 1.0), "mJy": ("flux", 1.0e-3), "uJy": ("flux", 1.0e-6), "nJy": ("flux", 1.0e-9),
    "Angstrom": ("length", 1.0), "nm": ("length", 10.0), "um": ("length", 1.0e4),
    "cm": ("length", 1.0e8), "Mpc": ("length", 3.0856775814913673e32),
    "dimensionless": ("none", 1.0),
}
_ALIASES = {"µJy": "uJy", "AA": "Angstrom", "angstrom": "Angstrom", "micron": "um",
            "1": "dimensionless", "": "dimensionless", "Msol": "Msun"}


class Unit:
    """A named unit with a dimension and a scale factor to the base unit."""

    __array_priority__ = 1000

    def __init__(self, name):
        if isinstance(name, Unit):
            name = name.name
        name = str(name)
        name = _ALIASES.get(name, name)
        if name not in _REGISTRY:
            raise ValueError(f"Unknown unit '{name}'")
        self.name = name
        self.dimensions, self.factor = _REGISTRY[name]

    # value * Unit  /  Unit * value
    def __rmul__(self, other):
        return Quantity(other, self)

    def __mul__(self, other):
        return Quantity(other, self)

    def __eq__(self, other):
        try:
            return Unit(str(other)).name == self.name
        except ValueError:
            return False

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return self.name

    __str__ = __repr__


class Quantity(np.ndarray):
    """ndarray carrying a :class:`Unit` (mirrors the used part of ``unyt_array``)."""

    __array_priority__ = 2000

    def __new__(cls, value, units="dimensionless"):
        if isinstance(value, Quantity) and units is None:
            units = value.units
        arr = np.asarray(value)
        if arr.dtype.kind != "f":
            arr = arr.astype(float)
        obj = arr.view(cls)
        obj.units = units if isinstance(units, Unit) else Unit(units)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.units = getattr(obj, "units", Unit("dimensionless"))

    @property
    def value(self):
        return np.asarray(self).view(np.ndarray).copy() if self.ndim else np.float64(np.asarray(self))

    @property
    def v(self):
        return self.value

    def to_value(self, units=None):
        if units is None:
            return self.value
        return self.to(units).value

    def to(self, units):
        units = units if isinstance(units, Unit) else Unit(units)
        if units.dimensions != self.units.dimensions:
            raise ValueError(f"Cannot convert {self.units} to {units}")
        if units.name == self.units.name:
            return Quantity(np.asarray(self), units)
        return Quantity(np.asarray(self) * (self.units.factor / units.factor), units)

    def __reduce__(self):
        state = super().__reduce__()
        return (state[0], state[1], state[2] + (self.units.name,))

    def __setstate__(self, state):
        self.units = Unit(state[-1])
        super().__setstate__(state[:-1])


def unyt_array(value, units="dimensionless"):
    return Quantity(value, units)


def unyt_quantity(value, units="dimensionless"):
    return Quantity(value, units)


def has_units(x) -> bool:
    return hasattr(x, "units") and hasattr(x, "value")


def as_quantity(x, default_units):
    """Return ``x`` as a Quantity; bare numbers get ``default_units``."""
    if isinstance(x, Quantity):
        return x
    if has_units(x):  # a real unyt object
        return Quantity(np.asarray(x.value, dtype=float), str(x.units))
    return Quantity(x, default_units)


def strip_units(x, units=None):
    """``x`` as a plain float ndarray, converted to ``units`` when it carries any."""
    if has_units(x):
        q = as_quantity(x, units or "dimensionless")
        return np.asarray(q.to(units).value if units is not None else q.value, dtype=float)
    return np.asarray(x, dtype=float)


yr, Myr, Gyr = Unit("yr"), Unit("Myr"), Unit("Gyr")
Msun = Unit("Msun")
Jy, mJy, uJy, nJy = Unit("Jy"), Unit("mJy"), Unit("uJy"), Unit("nJy")
Angstrom, um, nm, cm, Mpc = Unit("Angstrom"), Unit("um"), Unit("nm"), Unit("cm"), Unit("Mpc")
dimensionless = Unit("dimensionless")
