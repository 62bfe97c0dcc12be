"""Generate the committed golden fixtures under ``tests/golden/`` (TEST INFRASTRUCTURE ONLY).

Run in the build container (where ``/root/reference`` exists):

    python oracle/make_golden.py

It executes the reference's OWN, UNMODIFIED modules:

* ``zodipy/line_of_sight.py``, ``brightness.py``, ``number_density.py``, ``scattering.py``,
  ``component.py``, ``component_params.py`` through ``oracle/ref_loader.py``;
* ``zodipy/source_params.py``, ``zodiacal_light_model.py``, ``model_registry.py`` and
  ``unpack_model.py`` through a ~60-line stand-in for the handful of ``astropy.units`` features
  they touch (Astropy is not installed in this image) - so the per-model spectral scalars
  (emissivity, albedo, C1-C3, solar irradiance, calibration) in the fixtures come from the
  reference's own interpolation code and SciPy;
* drives them exactly like ``zodipy/model.py:253-279``.

The only reference input that cannot be produced by reference code here is the blackbody table
(``zodipy/blackbody.py:33-49`` needs ``astropy.modeling``); it is built from the published Planck
formula in ``oracle/zodi_oracle.blackbody_table`` and stored in the fixture, so the CUDA path and
the oracle are compared on the identical table.

Outputs: ``tests/golden/cases.npz`` (arrays), ``tests/golden/cases.json`` (model specs and case
descriptions), ``tests/golden/reference_tables.json`` (the reference's component / source tables,
for checking the product's carried-over data), ``tests/golden/dirbe_tabulated.json`` (values of the
reference's ``tests/dirbe_tabulated.py``).
"""
from __future__ import annotations

import dataclasses
import importlib
import inspect
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import zodi_oracle as oracle  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
C_LIGHT = 299792458.0


# ---------------------------------------------------------------------------------------
# Minimal astropy.units stand-in (only what source_params / unpack_model / zodiacal_light_model
# touch: Quantity(value, unit), .value, .unit, .to(unit, equivalencies), .to_value, unit division)
# ---------------------------------------------------------------------------------------
class Unit:
    def __init__(self, name, kind, scale):
        self.name, self.kind, self.scale = name, kind, scale

    def __truediv__(self, other):
        return Unit(f"{self.name}/{other.name}", f"{self.kind}/{other.kind}", self.scale / other.scale)

    def __rmul__(self, value):
        return Quantity(value, self)

    def __repr__(self):
        return self.name


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
        if unit.kind == self.unit.kind:
            return Quantity(np.asarray(self) * (self.unit.scale / unit.scale), unit)
        if {unit.kind, self.unit.kind} == {"length", "frequency"} and equivalencies == "spectral":
            si = np.asarray(self) * self.unit.scale  # m or Hz
            return Quantity((C_LIGHT / si) / unit.scale, unit)
        raise ValueError(f"cannot convert {self.unit} to {unit}")

    def to_value(self, unit, equivalencies=None):
        return self.to(unit, equivalencies).value


def install_astropy_shim():
    units = types.ModuleType("astropy.units")
    units.Quantity = Quantity
    units.micron = units.um = Unit("micron", "length", 1e-6)
    units.GHz = Unit("GHz", "frequency", 1e9)
    units.MJy = Unit("MJy", "sfd", 1e6)
    units.Jy = Unit("Jy", "sfd", 1.0)
    units.AU = Unit("AU", "length", 1.495978707e13)  # in cm for the RRM calibration ratio
    units.cm = Unit("cm", "length", 1.0)
    units.spectral = lambda: "spectral"
    units.UnitConversionError = ValueError
    astropy = types.ModuleType("astropy")
    astropy.units = units
    sys.modules["astropy"] = astropy
    sys.modules["astropy.units"] = units
    return units


def load_reference_full():
    """Hot-path modules + the real model/registry/unpack modules over the units stand-in."""
    units = install_astropy_shim()
    ns = ref_loader.load()
    # replace the two empty stubs by the real modules (they only need astropy.units)
    for name in ("zodiacal_light_model", "model_registry"):
        del sys.modules[f"zodipy.{name}"]
    for name in ("source_params", "zodiacal_light_model", "model_registry", "unpack_model"):
        setattr(ns, name, importlib.import_module(f"zodipy.{name}"))
    ns.units = units
    return ns


# ---------------------------------------------------------------------------------------
# Build a neutral model spec from reference objects (binding rule: number_density.py:441-463)
# ---------------------------------------------------------------------------------------
TYPE_NAMES = {
    "Cloud": "cloud", "Band": "band", "Ring": "ring", "Feature": "feature", "Fan": "fan",
    "Comet": "comet", "Interstellar": "interstellar", "NarrowBand": "narrow_band",
    "BroadBand": "broad_band", "RingRRM": "ring_rrm", "FeatureRRM": "feature_rrm",
}


def bound_density_params(ns, comp):
    func = ns.number_density.DENSITY_FUNCS[type(comp)]
    wanted = inspect.signature(func).parameters.keys()
    out = {}
    for k, v in dataclasses.asdict(comp).items():
        if k in wanted:
            out[k] = np.asarray(v, dtype=np.float64).reshape(-1).tolist() if k == "X_0" else float(v)
    return out


def build_spec(ns, name, x, unit, weights=None, deg=50, extrapolate=False, comps_override=None):
    """Mirror of what ``Model.__init__`` prepares (``zodipy/model.py:69-108,281-301``)."""
    u = ns.units
    model = ns.model_registry.model_registry.get_model(name)
    comps = comps_override if comps_override is not None else model.comps
    xq = Quantity(x, getattr(u, unit))
    if weights is not None:
        w = np.asarray(weights, dtype=np.float64)
        norm_w = w / oracle._trapezoid(w, np.asarray(x, dtype=np.float64))  # model.py:94
    else:
        norm_w = None
    unpack = ns.unpack_model.get_model_interp_func(model)
    if comps_override is not None:
        model = dataclasses.replace(model, comps=comps)
    comp_params, shared = unpack(xq, None if norm_w is None else Quantity(norm_w, None), model,
                                 not extrapolate)
    freq_hz = xq.to(Unit("Hz", "frequency", 1.0), "spectral").value
    table = oracle.blackbody_table(freq_hz, norm_w, x_native=x)
    points, wts = np.polynomial.legendre.leggauss(deg)  # model.py:103
    kind = "kelsall" if type(model).__name__ == "Kelsall" else "rrm"
    spec = {"kind": kind, "name": name, "comps": [], "table": table, "points": points,
            "weights": wts}
    for label, comp in comps.items():
        entry = {
            "label": label.value,
            "type": TYPE_NAMES[type(comp).__name__],
            "cutoff": [float(v) for v in ns.line_of_sight.COMPONENT_CUTOFFS[label]],
            "params": bound_density_params(ns, comp),
        }
        entry.update({k: float(v) for k, v in comp_params[label].items()})
        spec["comps"].append(entry)
    spec.update({k: float(v) for k, v in shared.items()})
    return spec, model, comp_params, shared, comps


def reference_emission(ns, model, comps, comp_params, shared, table, points, weights, u, obs,
                       earth):
    """Drive the reference modules exactly as ``zodipy/model.py:253-279`` does."""
    import functools

    los = ns.line_of_sight
    start, stop = los.get_line_of_sight_range(components=comps.keys(), unit_vectors=u, obs_pos=obs)
    partials = ns.number_density.get_partial_number_density_func(comps=comps)
    partials = ns.number_density.update_partial_earth_pos(partials, earth_pos=earth)
    shared_partial = functools.partial(model.brightness_at_step_callable,
                                       bp_interpolation_table=table, **shared)
    shared_partial = functools.partial(shared_partial, X_obs=obs)
    emission = np.zeros((len(comps), u.shape[1]))
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for idx, label in enumerate(comps.keys()):
            f = functools.partial(shared_partial, u_los=u, start=start[label], stop=stop[label],
                                  number_density_func=partials[label], **comp_params[label])
            emission[idx] = los.integrate_leggauss(f, points, weights)
    return emission, start, stop


# ---------------------------------------------------------------------------------------
# Inputs
# ---------------------------------------------------------------------------------------
def lonlat_to_vec(lon_deg, lat_deg):
    lon, lat = np.radians(lon_deg), np.radians(lat_deg)
    return np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])


def random_dirs(rng, n):
    v = rng.normal(size=(3, n))
    v /= np.linalg.norm(v, axis=0)
    # edge cases: both poles, the +-x/+-y axes, and a sun-ward ray for an observer near (-0.18, 0.97)
    edge = np.array([[0, 0, 1.0], [0, 0, -1.0], [1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0],
                     [0, -1.0, 0]]).T
    sun = lonlat_to_vec(np.array([259.4, 280.5]), np.array([0.0, 0.001]))
    v[:, : edge.shape[1]] = edge
    v[:, edge.shape[1]: edge.shape[1] + 2] = sun
    return v


def spec_to_json(spec):
    out = {}
    for k, v in spec.items():
        out[k] = v.tolist() if isinstance(v, np.ndarray) else v
    return out


def main():
    ns = load_reference_full()
    rng = np.random.default_rng(20261017)
    arrays, cases = {}, []

    obs_fix = np.array([[-0.18], [0.967], [0.00002]])
    u4 = lonlat_to_vec(np.array([0.0, 90.0, 200.0, 259.4]), np.array([0.0, 45.0, -80.0, 0.0]))
    u_rand = random_dirs(rng, 1000)
    n_tod = 600
    u_tod = random_dirs(rng, n_tod)
    ang = np.linspace(1.7, 1.7 + 2 * np.pi * 0.8, n_tod)
    earth_tod = np.array([0.9833 * np.cos(ang) * (1 + 0.0167 * np.cos(ang)),
                          0.9833 * np.sin(ang) * (1 + 0.0167 * np.cos(ang)),
                          1e-5 * np.sin(3 * ang)])
    obs_tod = earth_tod * 1.01 + np.array([[0.0], [0.0], [0.002]]) * np.cos(ang)

    dirbe_bp_x = np.linspace(9.0, 15.0, 10)
    dirbe_bp_w = np.exp(-0.5 * ((dirbe_bp_x - 12.0) / 1.5) ** 2)
    planck_bp_x = np.linspace(700.0, 857.0, 12)
    planck_bp_w = 1.0 + 0.3 * np.sin(np.linspace(0, 3, 12))

    def add(case_id, name, x, unit, u, obs, earth, weights=None, deg=50, extrapolate=False,
            mutate=None, note=""):
        comps_override = None
        if mutate is not None:
            base = ns.model_registry.model_registry.get_model(name).comps
            comps_override = {k: (mutate(k.value, v) or v) for k, v in base.items()}
        spec, model, comp_params, shared, comps = build_spec(
            ns, name, x, unit, weights=weights, deg=deg, extrapolate=extrapolate,
            comps_override=comps_override)
        em, start, stop = reference_emission(ns, model, comps, comp_params, shared, spec["table"],
                                             spec["points"], spec["weights"], u, obs, earth)
        arrays[f"{case_id}/u"] = u
        arrays[f"{case_id}/obs"] = obs
        arrays[f"{case_id}/earth"] = earth
        arrays[f"{case_id}/emission"] = em
        arrays[f"{case_id}/start"] = np.array([np.broadcast_to(start[k], (u.shape[1],)) for k in comps])
        arrays[f"{case_id}/stop"] = np.array([np.broadcast_to(stop[k], (u.shape[1],)) for k in comps])
        cases.append({"id": case_id, "model": name, "x": np.asarray(x).tolist(), "unit": unit,
                      "weights": None if weights is None else np.asarray(weights).tolist(),
                      "deg": deg, "extrapolate": extrapolate, "mutated": mutate is not None,
                      "note": note, "spec": spec_to_json(spec)})
        tot = em.sum(axis=0)
        print(f"{case_id:28s} N={u.shape[1]:5d} ncomps={em.shape[0]} total[min,max]="
              f"[{np.nanmin(tot):.6g}, {np.nanmax(tot):.6g}] nan={int(np.isnan(tot).sum())}")

    # SURVEY Appendix D fixed-input goldens G1-G4
    add("g1_dirbe25_fix", "dirbe", 25.0, "micron", u4, obs_fix, obs_fix, note="Appendix D G1")
    add("g2_dirbe1p25_fix", "dirbe", 1.25, "micron", u4, obs_fix, obs_fix, note="Appendix D G2")
    add("g3_planck18_857_fix", "planck18", 857.0, "GHz", u4, obs_fix, obs_fix, note="Appendix D G3")
    add("g4_rrm25_fix", "rrm-experimental", 25.0, "micron", u4, obs_fix, obs_fix, note="Appendix D G4")

    # every DIRBE band centre on random directions (thermal + scattering branches)
    for lam in (1.25, 2.2, 3.5, 4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0):
        add(f"dirbe_{str(lam).replace('.', 'p')}um", "dirbe", lam, "micron", u_rand[:, :400], obs_fix,
            obs_fix)
    add("dirbe_25um_rand", "dirbe", 25.0, "micron", u_rand, obs_fix, obs_fix,
        note="BASELINE config 1 physics")
    add("dirbe_12um_bandpass", "dirbe", dirbe_bp_x, "micron", u_rand, obs_fix, obs_fix,
        weights=dirbe_bp_w, note="BASELINE config 2 physics (10-sample bandpass)")
    add("dirbe_3um_interp", "dirbe", 3.0, "micron", u_rand[:, :400], obs_fix, obs_fix,
        note="albedo/emissivity linearly interpolated, C1-3 nearest")
    add("dirbe_2p85um_nearest_tie", "dirbe", 2.85, "micron", u_rand[:, :100], obs_fix, obs_fix,
        note="exact midpoint of 2.2/3.5 um: interp1d nearest tie rule")
    add("dirbe_300um_extrap", "dirbe", 300.0, "micron", u_rand[:, :200], obs_fix, obs_fix,
        extrapolate=True, note="linear extrapolation of spectral parameters")
    add("planck13_545", "planck13", 545.0, "GHz", u_rand, obs_fix, obs_fix,
        note="BASELINE config 5 physics (negative feature emissivity)")
    add("planck15_353", "planck15", 353.0, "GHz", u_rand[:, :400], obs_fix, obs_fix)
    add("planck18_857", "planck18", 857.0, "GHz", u_rand, obs_fix, obs_fix,
        note="BASELINE config 3 physics")
    add("planck18_bandpass", "planck18", planck_bp_x, "GHz", u_rand[:, :400], obs_fix, obs_fix,
        weights=planck_bp_w)
    add("odegard_217", "odegard", 217.0, "GHz", u_rand[:, :400], obs_fix, obs_fix)
    add("rrm_60um", "rrm-experimental", 60.0, "micron", u_rand, obs_fix, obs_fix)
    add("rrm_12um", "rrm-experimental", 12.0, "micron", u_rand[:, :400], obs_fix, obs_fix)

    # time-ordered: per-sample observer != earth (BASELINE config 4 physics)
    add("dirbe_25um_tod", "dirbe", 25.0, "micron", u_tod, obs_tod, earth_tod, note="TOD, obs != earth")
    add("dirbe_1p25um_tod", "dirbe", 1.25, "micron", u_tod[:, :300], obs_tod[:, :300],
        earth_tod[:, :300], note="TOD with scattering")
    add("rrm_25um_tod", "rrm-experimental", 25.0, "micron", u_tod[:, :300], obs_tod[:, :300],
        earth_tod[:, :300])
    add("planck18_tod", "planck18", 857.0, "GHz", u_tod[:, :300], obs_tod[:, :300], earth_tod[:, :300])

    # observers that trigger / avoid the .any() early-out (Q1)
    obs_far = np.array([[1.2], [0.9], [0.05]])  # r = 1.5008 > ring/feature outer cutoffs
    obs_in = np.array([[0.5], [-0.55], [0.01]])  # r = 0.743 < ring inner 0.8 -> start computed
    obs_mars = np.array([[-1.1], [1.2], [0.03]])  # r = 1.628 > R_MARS: rrm fan outside
    add("dirbe_25um_obs1p5", "dirbe", 25.0, "micron", u_rand[:, :400], obs_far, obs_fix)
    add("dirbe_25um_obs0p74", "dirbe", 25.0, "micron", u_rand[:, :400], obs_in, obs_fix)
    add("dirbe_2p2um_obs0p74", "dirbe", 2.2, "micron", u_rand[:, :200], obs_in, obs_fix)
    add("rrm_25um_obs1p63", "rrm-experimental", 25.0, "micron", u_rand[:, :400], obs_mars, obs_fix)
    obs_tod_straddle = obs_tod[:, :300] * np.linspace(0.75, 1.35, 300)  # crosses 0.8 / 1.2 / 1.3
    add("dirbe_25um_tod_straddle", "dirbe", 25.0, "micron", u_tod[:, :300], obs_tod_straddle,
        earth_tod[:, :300], note="global any() early-out with mixed observers")

    # quadrature degree and user-updated parameters
    add("dirbe_25um_deg20", "dirbe", 25.0, "micron", u_rand[:, :300], obs_fix, obs_fix, deg=20)
    add("dirbe_25um_deg100", "dirbe", 25.0, "micron", u_rand[:, :300], obs_fix, obs_fix, deg=100)
    add("planck18_deg7", "planck18", 857.0, "GHz", u_rand[:, :300], obs_fix, obs_fix, deg=7)

    def mutate(label, comp):
        if label == "band2":
            return dataclasses.replace(comp, p=3.7, x_0=0.01, y_0=-0.004, z_0=0.002)
        if label == "cloud":
            return dataclasses.replace(comp, gamma=1.1, mu=0.25)
        return None

    add("dirbe_25um_mutated", "dirbe", 25.0, "micron", u_rand[:, :400], obs_fix, obs_fix,
        mutate=mutate, note="update_parameters-style: band p != 4, band offset, cloud shape")

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cases.npz"), **arrays)
    with open(os.path.join(GOLDEN_DIR, "cases.json"), "w") as fh:
        json.dump({"generator": "oracle/make_golden.py", "numpy": np.__version__, "cases": cases}, fh)

    # the reference's parameter tables, for checking the product's carried-over data
    sp = ns.source_params
    tables = {"comps": {}, "source": {}}
    for set_name in ("DIRBE", "PLANCK", "RRM"):
        tables["comps"][set_name] = {
            k.value: {"type": TYPE_NAMES[type(v).__name__],
                      "fields": {f.name: getattr(v, f.name) for f in dataclasses.fields(v) if f.init}}
            for k, v in getattr(ns.component_params, set_name).items()}
    for k, v in vars(sp).items():
        if k.startswith("_") or k in ("u", "ComponentLabel"):
            continue
        if isinstance(v, dict):
            tables["source"][k] = {kk.value: (list(vv) if isinstance(vv, (tuple, list)) else vv)
                                   for kk, vv in v.items()}
        elif isinstance(v, Quantity):
            tables["source"][k] = {"value": np.asarray(v).tolist(), "unit": v.unit.name}
        elif isinstance(v, (tuple, list)):
            tables["source"][k] = list(v)
        elif isinstance(v, (int, float)):
            tables["source"][k] = v
    tables["cutoffs"] = {k.value: [float(a), float(b)]
                         for k, (a, b) in ns.line_of_sight.COMPONENT_CUTOFFS.items()}
    reg = ns.model_registry.model_registry
    tables["models"] = {name: {"class": type(reg.get_model(name)).__name__,
                               "comps": [k.value for k in reg.get_model(name).comps],
                               "spectrum": np.asarray(reg.get_model(name).spectrum).tolist(),
                               "spectrum_unit": reg.get_model(name).spectrum.unit.name}
                        for name in reg.models}
    with open(os.path.join(GOLDEN_DIR, "reference_tables.json"), "w") as fh:
        json.dump(tables, fh, indent=1)

    # values of the reference's tests/dirbe_tabulated.py (DIRBE IDL software output)
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, "tests", "dirbe_tabulated.py")).read()
    env = {}
    shim_time = types.SimpleNamespace(Time=lambda s: s)
    fake = types.ModuleType("astropy")
    fake.time, fake.units = shim_time, types.SimpleNamespace(micron=Unit("micron", "length", 1e-6), Quantity=Quantity)
    saved = sys.modules.get("astropy")
    sys.modules["astropy"] = fake
    try:
        exec(compile(src, "dirbe_tabulated.py", "exec"), env)
    finally:
        sys.modules["astropy"] = saved
    bp_x, bp_w = env["DIRBE_25um_BANDPASS"]
    with open(os.path.join(GOLDEN_DIR, "dirbe_tabulated.json"), "w") as fh:
        json.dump({"start_day": env["DIRBE_START_DAY"], "days": env["DAYS"], "lon": env["LON"],
                   "lat": env["LAT"],
                   "emission": {str(k): v for k, v in env["TABULATED_DIRBE_EMISSION"].items()},
                   "bandpass_25um_x": np.asarray(bp_x, dtype=float).tolist(),
                   "bandpass_25um_w": list(bp_w)}, fh, indent=1)
    print("wrote", GOLDEN_DIR)


if __name__ == "__main__":
    main()
