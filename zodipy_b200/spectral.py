"""Host-side spectral preparation: blackbody table and per-band model scalars (Astropy-free).

Produces the inputs the kernel consumes, following the reference's ``Model.__init__`` path:

* ``tabulate_blackbody_emission`` - the (2, 100) clamped-linear-interpolation table of
  ``zodipy/blackbody.py:9-13,33-49``.  Astropy's ``BlackBody`` model is the Planck law
  ``B_nu = 2 h nu^3 / c^2 / expm1(h nu / k T)`` in SI, converted to MJy/sr (x 1e20).
* ``interp_spectral_param`` - ``zodipy/unpack_model.py:140-173`` (SciPy ``interp1d`` linear /
  nearest, optional linear extrapolation, optional bandpass integration).
* ``unpack_model`` - ``zodipy/unpack_model.py:23-137``.

This is O(bandpass length) work done once per ``Model``; it stays on the host by design.
"""
from __future__ import annotations

import numpy as np

from . import units as zu

H_PLANCK = 6.62607015e-34  # J s      (CODATA 2018, exact)
K_BOLTZ = 1.380649e-23  # J / K    (CODATA 2018, exact)
N_TEMPS = 100
MIN_TEMP = 40.0
MAX_TEMP = 550.0
# units.Quantity(c, MJy/AU).to_value(Jy/cm), unpack_model.py:135-136
MJY_PER_AU_TO_JY_PER_CM = 1e6 / 1.495978707e13


def trapezoid(y, x):
    """Composite trapezoidal rule along the last axis (``scipy.integrate.trapezoid``)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return np.sum(np.diff(x) * (y[..., 1:] + y[..., :-1]) / 2.0, axis=-1)


def planck_mjy_sr(freq_hz, temps):
    """Planck specific intensity [MJy/sr] on the outer grid (len(freq), len(temps))."""
    nu = np.atleast_1d(np.asarray(freq_hz, dtype=np.float64))[:, None]
    t = np.asarray(temps, dtype=np.float64)[None, :]
    with np.errstate(over="ignore"):
        return 1e20 * (2.0 * H_PLANCK * nu**3 / zu.C_LIGHT**2) / np.expm1(H_PLANCK * nu / (K_BOLTZ * t))


def tabulate_blackbody_emission(x, normalized_weights=None) -> np.ndarray:
    """(2, N_TEMPS) table: temperatures [K] and (bandpass-integrated) B_nu [MJy/sr]."""
    temps = np.linspace(MIN_TEMP, MAX_TEMP, N_TEMPS)
    bnu = planck_mjy_sr(zu.spectral_value(x, "Hz"), temps)
    if normalized_weights is None:
        emission = bnu[0]
    else:
        # integral over the USER's x values, in the user's unit and order (blackbody.py:41-43)
        w = np.asarray(normalized_weights, dtype=np.float64)
        emission = trapezoid(w[None, :] * bnu.T, zu.native_value(x))
    return np.asarray([temps, emission])


def interp_spectral_param(x_model_unit, normalized_weights, spectrum, parameter,
                          use_nearest: bool = False, bounds_error: bool = True):
    """One spectral parameter at the requested x (already in the model spectrum's unit)."""
    knots = np.asarray(spectrum, dtype=np.float64)
    vals = np.asarray(parameter, dtype=np.float64)
    if knots[0] > knots[-1]:
        knots, vals = knots[::-1], vals[::-1]
    xq = np.asarray(x_model_unit, dtype=np.float64)
    if bounds_error and (np.any(xq < knots[0]) or np.any(xq > knots[-1])):
        raise ValueError("A value in x is outside the tabulated spectrum of the model.")
    if use_nearest:
        # interp1d(kind="nearest"): ties at a midpoint go to the lower knot
        midpoints = 0.5 * (knots[1:] + knots[:-1])
        out = vals[np.searchsorted(midpoints, xq, side="left")]
    else:
        hi = np.clip(np.searchsorted(knots, xq, side="left"), 1, knots.size - 1)
        lo = hi - 1
        out = (vals[hi] - vals[lo]) / (knots[hi] - knots[lo]) * (xq - knots[lo]) + vals[lo]
    if normalized_weights is not None:
        return float(trapezoid(np.asarray(normalized_weights) * out, xq))
    return float(out) if np.ndim(out) == 0 else out


def unpack_model(model, x, normalized_weights, bounds_error: bool):
    """Per-component and shared source parameters at x (``unpack_model.py:23-137``)."""
    spectrum = zu.native_value(model.spectrum)
    xv = zu.spectral_value(x, zu.unit_name(model.spectrum))

    def at(param, nearest=False):
        return interp_spectral_param(xv, normalized_weights, spectrum, param, nearest, bounds_error)

    comp_params: dict = {}
    shared: dict = {}
    if model.kind == "kelsall":
        shared["T_0"] = model.T_0
        shared["delta"] = model.delta
        for label in model.comps:
            comp_params[label] = {
                "emissivity": at(model.emissivities[label]),
                "albedo": at(model.albedos[label]) if model.albedos is not None else 0,
            }
        for name in ("C1", "C2", "C3"):
            table = getattr(model, name)
            shared[name] = at(table, nearest=True) if table is not None else 0
        shared["solar_irradiance"] = (
            at(model.solar_irradiance) if model.solar_irradiance is not None else 0)
    elif model.kind == "rrm":
        for label in model.comps:
            comp_params[label] = {"T_0": model.T_0[label], "delta": model.delta[label]}
        shared["calibration"] = at(model.calibration) * MJY_PER_AU_TO_JY_PER_CM
    else:
        raise TypeError(f"unknown model kind {model.kind!r}")
    return comp_params, shared
