"""CPU oracle for ZodiPy's line-of-sight brightness integration (TEST INFRASTRUCTURE ONLY).

This module is a NumPy restatement of the reference's array-only hot path
(``zodipy/model.py:253-279`` and everything it calls).  It exists to CHECK the CUDA path; it is
not part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product (``zodipy_b200``) never
falls back to it: without the CUDA extension the product raises.

Parity status: PINNED.  ``oracle/make_golden.py`` (run in the build container) executes the
reference's own unmodified modules (``oracle/ref_loader.py``) on fixed inputs and commits inputs,
model specs and outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures at <=1e-12 relative and against the DIRBE IDL table
(``tests/dirbe_tabulated.py:3-156`` of the reference, copied values in
``tests/golden/dirbe_tabulated.json``) at the reference test's own 1 % tolerance.

Third-party arithmetic: the reference delegates ``exp``/``power``/``arctan2``/``arcsin``/
``arccos``/``mod``/``interp``/``leggauss`` to NumPy (pinned 1.26.4 in the reference's
``requirements.txt:37``; this image has 2.3.5) and the Planck function to Astropy 6.0.1
(``requirements.txt:4``, not installed here; restated in :func:`blackbody_table` from the
published formula B_nu = 2 h nu^3 / c^2 / expm1(h nu / k T) with CODATA-2018 constants, which is
what ``astropy.modeling.physical_models.BlackBody`` evaluates).

The "model spec" consumed here is a plain dict (no product classes):

    spec = {
      "kind": "kelsall" | "rrm",
      "comps": [ {"label": str, "type": str, "cutoff": (inner, outer), "params": {...},
                  # kelsall: "emissivity", "albedo"      rrm: "T_0", "delta"
                 }, ...],
      # kelsall: "T_0", "delta", "C1", "C2", "C3", "solar_irradiance"     rrm: "calibration"
      "table": ndarray (2, n_T),  "points": ndarray (deg,), "weights": ndarray (deg,),
    }

``params`` holds exactly the keyword arguments the reference binds from the component dataclass
(``zodipy/number_density.py:441-463``): ``X_0`` (3,), ``sin_Omega_rad``, ``cos_Omega_rad``,
``sin_i_rad``, ``cos_i_rad`` and the per-type fields.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # R_0, zodipy/line_of_sight.py:14

# CODATA-2018 (exact SI) constants used by astropy.constants in astropy 6.0.1
H_PLANCK = 6.62607015e-34
C_LIGHT = 299792458.0
K_BOLTZ = 1.380649e-23


# --------------------------------------------------------------------------------------
# Host-side inputs (restated without Astropy)
# --------------------------------------------------------------------------------------
def blackbody_table(freq_hz, norm_weights=None, x_native=None, n_temps=100, t_min=40.0,
                    t_max=550.0):
    """B_nu table, restating ``zodipy/blackbody.py:9-13,33-49``.

    ``freq_hz``: scalar or (m,) frequencies [Hz] (Astropy converts wavelengths with
    ``units.spectral()``: nu = c / lambda).  With a bandpass, the integral is
    ``trapezoid(w_norm * B(x_i, T), x)`` over the USER's native x values (``x_native``) in the
    user's order (``zodipy/blackbody.py:41-43``; quirk Q6).  Returns (2, n_temps): row 0
    temperatures [K], row 1 emission [MJy/sr].
    """
    temps = np.linspace(t_min, t_max, n_temps)  # blackbody.py:9-12
    nu = np.atleast_1d(np.asarray(freq_hz, dtype=np.float64))
    # BlackBody.evaluate: 2 h nu^3 / (c^2 expm1(h nu / k T)) [W m^-2 Hz^-1 sr^-1]; 1 MJy = 1e-20
    with np.errstate(over="ignore"):
        bnu = (2.0 * H_PLANCK * nu[:, None] ** 3 / C_LIGHT**2) / np.expm1(
            (H_PLANCK * nu[:, None]) / (K_BOLTZ * temps[None, :])
        )
    bnu *= 1e20
    if norm_weights is None:
        emission = bnu[0]
    else:
        w = np.asarray(norm_weights, dtype=np.float64)
        x = np.asarray(x_native, dtype=np.float64)
        emission = _trapezoid(w[None, :] * bnu.T, x)  # blackbody.py:41-43
    return np.asarray([temps, emission])


def _trapezoid(y, x):
    """``scipy.integrate.trapezoid`` along the last axis (same formula as NumPy's)."""
    d = np.diff(x)
    return np.sum(d * (y[..., 1:] + y[..., :-1]) / 2.0, axis=-1)


def interp_spectral_param(x_model_unit, norm_weights, spectrum, parameter, use_nearest=False,
                          bounds_error=True):
    """Restates ``zodipy/unpack_model.py:140-173`` (scipy ``interp1d`` linear / nearest).

    ``x_model_unit``: user's x converted to the model spectrum's unit (scalar or (m,)).
    """
    spectrum = np.asarray(spectrum, dtype=np.float64)
    parameter = np.asarray(parameter, dtype=np.float64)
    if not np.array_equal(spectrum, np.sort(spectrum)):  # unpack_model.py:151-153
        spectrum = spectrum[::-1]
        parameter = parameter[::-1]
    xq = np.asarray(x_model_unit, dtype=np.float64)
    if bounds_error and (np.any(xq < spectrum[0]) or np.any(xq > spectrum[-1])):
        raise ValueError("A value in x_new is outside the interpolation range.")
    if use_nearest:
        # scipy interp1d(kind="nearest"): index = searchsorted(midpoints, x, side="left"),
        # i.e. ties at a midpoint round DOWN to the lower knot; extrapolation -> end knots.
        mids = 0.5 * (spectrum[1:] + spectrum[:-1])
        idx = np.searchsorted(mids, xq, side="left")
        val = parameter[np.clip(idx, 0, len(parameter) - 1)]
    else:
        # linear with linear extrapolation from the end segments (fill_value="extrapolate")
        idx = np.clip(np.searchsorted(spectrum, xq, side="left"), 1, len(spectrum) - 1)
        lo, hi = idx - 1, idx
        slope = (parameter[hi] - parameter[lo]) / (spectrum[hi] - spectrum[lo])
        val = slope * (xq - spectrum[lo]) + parameter[lo]
    if norm_weights is not None:
        return _trapezoid(np.asarray(norm_weights) * val, np.asarray(xq))  # :171-172
    return val


def leggauss(deg):
    """Nodes/weights as built in ``zodipy/model.py:103``."""
    return np.polynomial.legendre.leggauss(deg)


# --------------------------------------------------------------------------------------
# Range: zodipy/line_of_sight.py:64-105
# --------------------------------------------------------------------------------------
def sphere_intersection(obs, u, cutoff):
    """Distance observer -> heliocentric sphere of radius ``cutoff`` along each ray.

    Follows ``zodipy/line_of_sight.py:64-85`` including its quirks: the GLOBAL ``.any()``
    early-out (Q1) and the missing ``z_0 * u_z`` term in ``b`` (Q2).
    """
    x0, y0, z0 = obs
    r_obs = np.sqrt(x0**2 + y0**2 + z0**2)
    if np.any(r_obs > cutoff):  # :72-73
        return np.full(obs.shape[-1], EPS)
    ux, uy, uz = u
    lon = np.arctan2(uy, ux)  # :76
    lat = np.arcsin(uz)  # :77
    cl = np.cos(lat)
    b = 2 * (x0 * cl * np.cos(lon) + y0 * cl * np.sin(lon))  # :80
    c = r_obs**2 - cutoff**2  # :81
    q = -0.5 * b * (1 + np.sqrt(b**2 - 4 * c) / np.abs(b))  # :83
    return np.maximum(q, c / q)  # :85


def los_range(spec, u, obs):
    """(start, stop) lists per component, ``zodipy/line_of_sight.py:88-105``."""
    start = [sphere_intersection(obs, u, c["cutoff"][0]) for c in spec["comps"]]
    stop = [sphere_intersection(obs, u, c["cutoff"][1]) for c in spec["comps"]]
    return start, stop


# --------------------------------------------------------------------------------------
# Densities: zodipy/number_density.py:47-404
# --------------------------------------------------------------------------------------
def _plane_geometry(X, p):
    """Common prologue of every density (e.g. ``number_density.py:61-67``): offset position,
    distance from the component centre and height above its symmetry plane.  Row-wise to avoid
    (3, N) temporaries (this port doubles as the CPU baseline and should not be slower than the
    reference's own code); same operations in the same order as the reference."""
    x0, y0, z0 = (float(v) for v in p["X_0"])
    xc, yc, zc = X[0] - x0, X[1] - y0, X[2] - z0
    Rc = np.sqrt(xc**2 + yc**2 + zc**2)
    Zc = (
        xc * p["sin_Omega_rad"] * p["sin_i_rad"]
        - yc * p["cos_Omega_rad"] * p["sin_i_rad"]
        + zc * p["cos_i_rad"]
    )
    return (xc, yc, zc), Rc, Zc


def density_cloud(X, p, earth=None):  # number_density.py:47-73
    _, Rc, Zc = _plane_geometry(X, p)
    zeta = np.abs(Zc / Rc)
    g = np.where(zeta < p["mu"], zeta**2 / (2 * p["mu"]), zeta - (p["mu"] / 2))
    return p["n_0"] * Rc ** -p["alpha"] * np.exp(-p["beta"] * g ** p["gamma"])


# Test aid.  The reference evaluates the band's radial cut-off literally as 1 - exp(-x) with
# x = (R/delta_r)**20 (number_density.py:108).  For x << 1 the result carries an ABSOLUTE rounding
# error of up to one ulp of 1.0 (2.2e-16) whatever exp implementation is used, i.e. a relative error
# eps / x that reaches 1e-6 for an observer at 0.45 AU - far above the 1e-10 parity target, on
# components that are then ~1e-9 of the total.  "unity" evaluates the band with the cut-off factor
# set to 1, which turns that per-node noise bound into an integral (reference_rounding_noise).
BAND_RADIAL_MODE = "literal"  # "literal" | "unity"


def density_band(X, p, earth=None):  # number_density.py:76-110
    _, Rc, Zc = _plane_geometry(X, p)
    zeta = np.abs(Zc / Rc)
    s = zeta / p["delta_zeta_rad"]
    t1 = 3 * p["n_0"] / Rc
    t2 = np.exp(-(s**6))
    t3 = 1 + (s ** p["p"]) / p["v"]
    t4 = 1.0 if BAND_RADIAL_MODE == "unity" else 1 - np.exp(-((Rc / p["delta_r"]) ** 20))
    return t1 * t2 * t3 * t4


def reference_rounding_noise(spec, u, obs, earth, ulps=2.0):
    """Bound (ncomps, N) on how much two correctly working implementations of the reference's
    formula may differ because of the literal 1 - exp(-x) of the band components (quirk Q9):
    `ulps` x 2.2e-16 x |integral of the band with its cut-off factor set to 1|; zero for the other
    component types."""
    global BAND_RADIAL_MODE
    BAND_RADIAL_MODE = "unity"
    try:
        unity = evaluate(spec, u, obs, earth)
    finally:
        BAND_RADIAL_MODE = "literal"
    is_band = np.array([c["type"] == "band" for c in spec["comps"]])[:, None]
    return np.where(is_band, ulps * EPS * np.abs(unity), 0.0)


def density_ring(X, p, earth=None):  # number_density.py:113-139
    _, Rc, Zc = _plane_geometry(X, p)
    t1 = -((Rc - p["R"]) ** 2) / p["sigma_r"] ** 2
    t2 = np.abs(Zc) / p["sigma_z"]
    return p["n_0"] * np.exp(t1 - t2)


def density_feature(X, p, earth=None):  # number_density.py:142-181
    Xc, Rc, Zc = _plane_geometry(X, p)
    Xe = earth - np.asarray(p["X_0"], dtype=np.float64).reshape(3, 1)
    theta = np.arctan2(Xc[1], Xc[0]) - np.arctan2(Xe[1], Xe[0])
    dth = theta - p["theta_rad"]
    dth = (dth + np.pi) % (2 * np.pi) - np.pi  # floored mod -> [-pi, pi)
    e = (Rc - p["R"]) ** 2 / p["sigma_r"] ** 2
    e = e + np.abs(Zc) / p["sigma_z"]
    e = e + dth**2 / p["sigma_theta_rad"] ** 2
    return p["n_0"] * np.exp(-e)


def _fan_like(X, p, inner, with_cos_q):
    """Shared body of fan (``number_density.py:184-218``) and comet (``:221-256``)."""
    Xc, Rc, Zc = _plane_geometry(X, p)
    inside = Rc <= p["R_outer"]
    if inner is not None:
        inside &= Rc >= inner
    out = np.zeros_like(Rc)
    with np.errstate(invalid="ignore", divide="ignore"):
        sin_beta = Zc / Rc
        beta = np.arcsin(sin_beta)
        za = np.abs(Zc)
        eps_ = np.where(za < p["Z_0"], 2 - (za / p["Z_0"]), 1)
        f = np.exp(-p["P"] * np.sin(np.abs(beta) ** eps_))
        if with_cos_q:
            f = np.cos(beta) ** p["Q"] * f
        val = (Rc ** (-p["gamma"])) * f
    out[inside] = val[inside]
    return out


def density_fan(X, p, earth=None):
    return _fan_like(X, p, None, True)


def density_comet(X, p, earth=None):
    return p["amp"] * _fan_like(X, p, p["R_inner"], False)


def density_interstellar(X, p, earth=None):  # number_density.py:259-264 (shape (1,), Q8)
    return np.array([p["amp"]])


def density_narrow_band(X, p, earth=None):  # number_density.py:267-304
    _, Rc, Zc = _plane_geometry(X, p)
    inside = (Rc >= p["R_inner"]) & (Rc <= p["R_outer"])
    out = np.zeros_like(Rc)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        beta_abs = np.abs(np.rad2deg(np.arcsin(Zc / Rc)))
        f = np.where(beta_abs < p["beta_nb"], np.exp(p["G"] * (beta_abs - p["beta_nb"])), 0)
        val = p["A"] * ((Rc / p["R_outer"]) ** (-p["gamma"])) * f
    out[inside] = val[inside]
    return out


def density_broad_band(X, p, earth=None):  # number_density.py:307-342
    _, Rc, Zc = _plane_geometry(X, p)
    inside = (Rc >= p["R_inner"]) & (Rc <= p["R_outer"])
    out = np.zeros_like(Rc)
    with np.errstate(invalid="ignore", divide="ignore"):
        beta = np.rad2deg(np.arcsin(Zc / Rc))
        f = np.exp(-0.5 * ((beta - p["beta_bb"]) / p["sigma_bb"]) ** 2) + np.exp(
            -0.5 * ((beta + p["beta_bb"]) / p["sigma_bb"]) ** 2
        )
        val = p["A"] * ((Rc / p["R_outer"]) ** (-p["gamma"])) * f
    out[inside] = val[inside]
    return out


def density_ring_rrm(X, p, earth=None):  # number_density.py:345-370
    return p["A"] * density_ring(X, p)


def density_feature_rrm(X, p, earth=None):  # number_density.py:373-404
    return p["A"] * density_feature(X, p, earth)


DENSITY = {  # number_density.py:408-420
    "cloud": density_cloud,
    "band": density_band,
    "ring": density_ring,
    "feature": density_feature,
    "fan": density_fan,
    "comet": density_comet,
    "interstellar": density_interstellar,
    "narrow_band": density_narrow_band,
    "broad_band": density_broad_band,
    "ring_rrm": density_ring_rrm,
    "feature_rrm": density_feature_rrm,
}


# --------------------------------------------------------------------------------------
# Source function: zodipy/brightness.py, zodipy/scattering.py, zodipy/blackbody.py:16-30
# --------------------------------------------------------------------------------------
def scattering_angle(R_los, R_h, X_los, X_h):  # scattering.py:11-31
    ct = (X_los * X_h).sum(axis=0) / (R_los * R_h)
    return np.arccos(-np.clip(ct, -1, 1))


def phase_function(theta, C1, C2, C3):  # scattering.py:34-59
    norm = 1 / (2 * np.pi * (2 * C1 + np.pi * C2 + (np.exp(C3 * np.pi) + 1) / (C3**2 + 1)))
    return norm * (C1 + C2 * theta + np.exp(C3 * theta))


def _geometry_at_node(r, start, stop, obs, u):
    R_los = 0.5 * (stop - start) * r + 0.5 * (stop + start)  # brightness.py:41
    X_los = R_los * u
    X_h = X_los + obs
    R_h = np.sqrt(X_h[0] ** 2 + X_h[1] ** 2 + X_h[2] ** 2)
    return R_los, X_los, X_h, R_h


def kelsall_step(r, start, stop, obs, u, earth, spec, comp):  # brightness.py:21-56
    R_los, X_los, X_h, R_h = _geometry_at_node(r, start, stop, obs, u)
    T = spec["T_0"] * R_h ** -spec["delta"]  # blackbody.py:30
    B = np.interp(T, spec["table"][0], spec["table"][1])  # brightness.py:48
    em = (1 - comp["albedo"]) * (comp["emissivity"] * B)
    if comp["albedo"] != 0:  # :50
        flux = spec["solar_irradiance"] / R_h**2
        theta = scattering_angle(R_los, R_h, X_los, X_h)
        em = em + comp["albedo"] * flux * phase_function(theta, spec["C1"], spec["C2"], spec["C3"])
    n = DENSITY[comp["type"]](X_h, comp["params"], earth)
    return em * n * 0.5 * (stop - start)


def rrm_step(r, start, stop, obs, u, earth, spec, comp):  # brightness.py:59-83
    _, _, X_h, R_h = _geometry_at_node(r, start, stop, obs, u)
    T = comp["T_0"] * R_h ** -comp["delta"]
    B = np.interp(T, spec["table"][0], spec["table"][1])
    n = DENSITY[comp["type"]](X_h, comp["params"], earth)
    return B * n * spec["calibration"] * 0.5 * (stop - start)


def evaluate(spec, u, obs, earth):
    """Restates the driver loop ``zodipy/model.py:253-279``.

    u: (3, N) ecliptic unit vectors; obs, earth: (3, 1) or (3, N) [AU].
    Returns emission (ncomps, N) float64 [MJy/sr].
    """
    u = np.asarray(u, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64).reshape(3, -1)
    earth = np.asarray(earth, dtype=np.float64).reshape(3, -1)
    n = u.shape[1]
    start, stop = los_range(spec, u, obs)
    step = kelsall_step if spec["kind"] == "kelsall" else rrm_step
    out = np.zeros((len(spec["comps"]), n))
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for ci, comp in enumerate(spec["comps"]):
            acc = 0  # integrate_leggauss, line_of_sight.py:61: sum() starts from int 0
            for x, w in zip(spec["points"], spec["weights"]):
                acc = acc + step(x, start[ci], stop[ci], obs, u, earth, spec, comp) * w
            out[ci] = acc
    return out


def outside_flags(spec, obs):
    """The 2*ncomps booleans of the ``.any()`` early-out (Q1) for a given observer array."""
    obs = np.asarray(obs, dtype=np.float64).reshape(3, -1)
    r = np.sqrt(obs[0] ** 2 + obs[1] ** 2 + obs[2] ** 2)
    return np.array([[bool(np.any(r > c["cutoff"][0])), bool(np.any(r > c["cutoff"][1]))]
                     for c in spec["comps"]], dtype=np.uint8)


# --------------------------------------------------------------------------------------
# CPU-baseline driver: the reference's nprocesses path, zodipy/model.py:182-198
# --------------------------------------------------------------------------------------
def _worker(args):
    spec, u, obs, earth = args
    return evaluate(spec, u, obs, earth)


def evaluate_parallel(spec, u, obs, earth, nprocesses):
    """``np.array_split`` + fork ``Pool`` + ``apply_async`` + concatenate (model.py:182-198)."""
    import multiprocessing

    u = np.asarray(u)
    obs = np.asarray(obs, dtype=np.float64).reshape(3, -1)
    earth = np.asarray(earth, dtype=np.float64).reshape(3, -1)
    n = u.shape[1]
    if not (n > nprocesses > 1):
        return evaluate(spec, u, obs, earth)
    u_s = np.array_split(u, nprocesses, axis=-1)
    obs_s = np.array_split(obs, nprocesses, axis=-1) if obs.shape[1] == n else [obs] * nprocesses
    ear_s = np.array_split(earth, nprocesses, axis=-1) if earth.shape[1] == n else [earth] * nprocesses
    with multiprocessing.get_context("fork").Pool(nprocesses) as pool:
        parts = [pool.apply_async(_worker, ((spec, a, b, c),)) for a, b, c in zip(u_s, obs_s, ear_s)]
        return np.concatenate([p.get() for p in parts], axis=-1)
