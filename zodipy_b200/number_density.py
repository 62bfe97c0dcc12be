"""``grid_number_density``: component densities tabulated on a Cartesian mesh.

Mirrors the reference utility ``zodipy/number_density.py:482-536`` (used for the density plots in
its documentation).  The mesh construction and the Earth ephemeris stay on the host; the densities
are evaluated by the library's device routines (``zodi_number_density``) - the same ones the
line-of-sight kernels use.
"""
from __future__ import annotations

import numpy as np

from . import units as zu
from .engine import DeviceModel
from .spec import build_spec
from .zodiacal_light_model import ZodiacalLightModel, clone_model, model_registry


def _length_au(v):
    return np.asarray(zu.length_value(v, "AU") if zu.is_quantity(v) else v, dtype=np.float64)


def grid_number_density_xyz(x, y, z, earth_xyz, model="dirbe", device: int = 0) -> np.ndarray:
    """Densities on ``np.meshgrid(x, y, z)`` for a given Earth position [AU] (no Astropy needed).

    Returns an array of shape (ncomps, len(y), len(x), len(z)) like the reference
    (``np.meshgrid`` default "xy" indexing, ``number_density.py:510,528-531``).
    """
    if isinstance(model, str):
        ipd_model = clone_model(model_registry.get_model(model))
    elif isinstance(model, ZodiacalLightModel):
        ipd_model = model
    else:
        raise TypeError("model type must be a `str` or a `ZodiacalLightModel`.")
    grid = np.asarray(np.meshgrid(_length_au(x), _length_au(y), _length_au(z)))
    # densities do not depend on wavelength: any in-range x builds the component block
    spectrum = zu.native_value(ipd_model.spectrum)
    spec = build_spec(ipd_model, zu.Quantity(float(spectrum[0]), zu.unit_name(ipd_model.spectrum)), None, True, 2)
    dm = DeviceModel(spec, device)
    try:
        dens = dm.number_density(grid.reshape(3, -1), earth_xyz)
    finally:
        dm.close()
    return dens.reshape((len(spec["comps"]), *grid.shape[1:]))


def grid_number_density(x, y, z, obstime, model="dirbe", ephemeris: str = "builtin", device: int = 0):
    """Reference signature (``number_density.py:482-489``); needs Astropy for the Earth position."""
    if not isinstance(model, (str, ZodiacalLightModel)):
        raise TypeError("model type must be a `str` or a `ZodiacalLightModel`.")
    from . import astro

    earth = astro._body_xyz("earth", obstime, ephemeris).flatten()
    return grid_number_density_xyz(x, y, z, earth, model, device)
