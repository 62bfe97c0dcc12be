"""Model specification handed to the device: neutral dict form and C-ABI descriptor.

``build_spec`` collects what the reference's ``Model.__init__`` prepares for the hot path
(``zodipy/model.py:89-108,281-301``): per-component density parameters bound by name
(``zodipy/number_density.py:441-463``), cutoff radii (``zodipy/line_of_sight.py:19-52``),
interpolated source scalars, the blackbody table and the Gauss-Legendre rule.  The dict layout is
documented in ``oracle/zodi_oracle.py`` (the oracle consumes the same dict, which is what makes
the parity tests compare like with like).  ``pack_desc`` lowers it to ``zodi_model_desc``.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from . import model_data as md
from . import spectral
from .component import COMPONENT_CLASSES


def build_spec(model, x, normalized_weights, bounds_error: bool, gauss_quad_degree: int) -> dict:
    comp_params, shared = spectral.unpack_model(model, x, normalized_weights, bounds_error)
    table = spectral.tabulate_blackbody_emission(x, normalized_weights)
    points, weights = np.polynomial.legendre.leggauss(gauss_quad_degree)  # model.py:103
    spec = {"kind": model.kind, "comps": [], "table": table, "points": points, "weights": weights}
    for label, comp in model.comps.items():
        if label.value not in md.COMPONENT_CUTOFFS:
            raise KeyError(f"no line-of-sight cutoff registered for component {label.value!r}")
        entry = {
            "label": label.value,
            "type": comp.type_tag,
            "cutoff": [float(v) for v in md.COMPONENT_CUTOFFS[label.value]],
            "params": comp.density_params(),
        }
        entry.update({k: float(v) for k, v in comp_params[label].items()})
        spec["comps"].append(entry)
    spec.update({k: float(v) for k, v in shared.items()})
    return spec


def pack_desc(spec: dict):
    """dict spec -> (ModelDesc, keepalive arrays).  The arrays must outlive the C call."""
    d = _cabi.ModelDesc()
    d.abi_version = _cabi.ABI_VERSION
    d.kind = _cabi.KELSALL if spec["kind"] == "kelsall" else _cabi.RRM
    comps = spec["comps"]
    if not 1 <= len(comps) <= _cabi.MAX_COMPS:
        raise ValueError(f"number of components {len(comps)} outside [1, {_cabi.MAX_COMPS}]")
    temps = np.ascontiguousarray(spec["table"][0], dtype=np.float64)
    bnu = np.ascontiguousarray(spec["table"][1], dtype=np.float64)
    nodes = np.ascontiguousarray(spec["points"], dtype=np.float64)
    weights = np.ascontiguousarray(spec["weights"], dtype=np.float64)
    d.n_comps, d.n_nodes, d.n_temps = len(comps), nodes.size, temps.size
    for name in ("T_0", "delta", "C1", "C2", "C3", "solar_irradiance", "calibration"):
        setattr(d, name, float(spec.get(name, 0.0)))
    d.temps, d.bnu = _cabi.as_double_p(temps), _cabi.as_double_p(bnu)
    d.nodes, d.weights = _cabi.as_double_p(nodes), _cabi.as_double_p(weights)
    for i, c in enumerate(comps):
        cd = d.comps[i]
        cls = COMPONENT_CLASSES[c["type"]]
        p = c["params"]
        cd.type = cls.type_id
        x0 = p.get("X_0", [0.0, 0.0, 0.0])
        for k in range(3):
            cd.x0[k] = float(x0[k])
        cd.sin_Omega = float(p.get("sin_Omega_rad", 0.0))
        cd.cos_Omega = float(p.get("cos_Omega_rad", 1.0))
        cd.sin_i = float(p.get("sin_i_rad", 0.0))
        cd.cos_i = float(p.get("cos_i_rad", 1.0))
        for k, name in enumerate(cls.shape_layout):
            cd.shape[k] = float(p[name])
        cd.cutoff_inner, cd.cutoff_outer = float(c["cutoff"][0]), float(c["cutoff"][1])
        cd.emissivity = float(c.get("emissivity", 0.0))
        cd.albedo = float(c.get("albedo", 0.0))
        cd.T_0 = float(c.get("T_0", spec.get("T_0", 0.0)))
        cd.delta = float(c.get("delta", spec.get("delta", 0.0)))
    return d, (temps, bnu, nodes, weights)


def outside_flags(spec: dict, r_max: float) -> np.ndarray:
    """(n_comps, 2) early-out flags for a largest observer distance (line_of_sight.py:72)."""
    return np.array([[r_max > c["cutoff"][0], r_max > c["cutoff"][1]] for c in spec["comps"]],
                    dtype=np.uint8)
