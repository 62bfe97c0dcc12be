"""Zodiacal component classes (host data model).

Same public names and constructor fields as the reference's ``zodipy/component.py:14-223`` so
that ``Model.get_parameters()`` / ``update_parameters()`` dictionaries are interchangeable, but
built from ONE declarative schema that also drives the C-ABI packing
(``include/zodi_b200.h: zodi_component_desc.shape[]``): each entry lists the constructor fields,
the derived ``*_rad`` fields (``component.py:85-89,130-136``) and the order in which the density
function's parameters (``zodipy/number_density.py:47-404``) are laid out in ``shape[]``.
"""
from __future__ import annotations

import dataclasses
from enum import Enum

import numpy as np

GEOMETRY_FIELDS = ("x_0", "y_0", "z_0", "i", "Omega")

# class name -> (type tag, C-ABI type id, constructor fields after the geometry,
#                {derived field: source field in degrees}, shape[] layout, needs Earth position)
SCHEMA = {
    "Cloud": ("cloud", 0, ("n_0", "alpha", "beta", "gamma", "mu"), {},
              ("n_0", "alpha", "beta", "gamma", "mu"), False),
    "Band": ("band", 1, ("n_0", "delta_zeta", "v", "p", "delta_r"), {"delta_zeta_rad": "delta_zeta"},
             ("n_0", "delta_zeta_rad", "v", "p", "delta_r"), False),
    "Ring": ("ring", 2, ("n_0", "R", "sigma_r", "sigma_z"), {}, ("n_0", "R", "sigma_r", "sigma_z"),
             False),
    "Feature": ("feature", 3, ("n_0", "R", "sigma_r", "sigma_z", "theta", "sigma_theta"),
                {"theta_rad": "theta", "sigma_theta_rad": "sigma_theta"},
                ("n_0", "R", "sigma_r", "sigma_z", "theta_rad", "sigma_theta_rad"), True),
    "Fan": ("fan", 4, ("gamma", "Z_0", "Q", "P", "R_outer"), {},
            ("Q", "P", "gamma", "Z_0", "R_outer"), False),
    "Comet": ("comet", 5, ("gamma", "Z_0", "P", "amp", "R_inner", "R_outer"), {},
              ("gamma", "Z_0", "P", "amp", "R_inner", "R_outer"), False),
    "Interstellar": ("interstellar", 6, ("amp",), {}, ("amp",), False),
    "NarrowBand": ("narrow_band", 7, ("gamma", "A", "G", "R_inner", "R_outer", "beta_nb"), {},
                   ("beta_nb", "G", "gamma", "A", "R_inner", "R_outer"), False),
    "BroadBand": ("broad_band", 8, ("gamma", "A", "R_inner", "R_outer", "beta_bb", "sigma_bb"), {},
                  ("beta_bb", "sigma_bb", "gamma", "A", "R_inner", "R_outer"), False),
    "RingRRM": ("ring_rrm", 9, ("n_0", "R", "sigma_r", "sigma_z", "A"), {},
                ("n_0", "R", "sigma_r", "sigma_z", "A"), False),
    "FeatureRRM": ("feature_rrm", 10, ("n_0", "R", "sigma_r", "sigma_z", "theta", "sigma_theta", "A"),
                   {"theta_rad": "theta", "sigma_theta_rad": "sigma_theta"},
                   ("n_0", "R", "sigma_r", "sigma_z", "theta_rad", "sigma_theta_rad", "A"), True),
}


class ComponentLabel(Enum):
    """Labels of the components of all shipped models (``zodipy/component.py:207-223``)."""

    CLOUD = "cloud"
    BAND1 = "band1"
    BAND2 = "band2"
    BAND3 = "band3"
    RING = "ring"
    FEATURE = "feature"
    FAN = "fan"
    COMET = "comet"
    INTERSTELLAR = "interstellar"
    INNER_NARROW_BAND = "inner_narrow_band"
    OUTER_NARROW_BAND = "outer_narrow_band"
    BROAD_BAND = "broad_band"
    RING_RRM = "ring_rrm"
    FEATURE_RRM = "feature_rrm"


class ZodiacalComponent:
    """Base of all component classes: geometry + derived trigonometric terms."""

    type_tag: str = ""
    type_id: int = -1
    shape_layout: tuple = ()
    needs_earth: bool = False
    _derived_deg: dict = {}

    def __post_init__(self) -> None:
        # derived quantities exactly as the reference forms them (component.py:39-44)
        self.X_0 = np.expand_dims([self.x_0, self.y_0, self.z_0], axis=-1)
        self.sin_i_rad = np.sin(np.radians(self.i))
        self.cos_i_rad = np.cos(np.radians(self.i))
        self.sin_Omega_rad = np.sin(np.radians(self.Omega))
        self.cos_Omega_rad = np.cos(np.radians(self.Omega))
        for name, src in self._derived_deg.items():
            setattr(self, name, np.radians(getattr(self, src)))

    def init_fields(self) -> dict:
        """Constructor fields only (what ``to_dict`` exports, zodiacal_light_model.py:44-49)."""
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.init}

    def shape_params(self) -> list:
        """Density-function parameters in ``zodi_component_desc.shape[]`` order."""
        return [float(getattr(self, name)) for name in self.shape_layout]

    def density_params(self) -> dict:
        """Keyword arguments the reference binds to the density function
        (``number_density.py:441-463``): geometry terms + the type's own fields."""
        out = {"X_0": [float(self.x_0), float(self.y_0), float(self.z_0)]}
        if self.type_tag != "interstellar":
            for k in ("sin_Omega_rad", "cos_Omega_rad", "sin_i_rad", "cos_i_rad"):
                out[k] = float(getattr(self, k))
        else:
            out = {}
        for name in self.shape_layout:
            out[name] = float(getattr(self, name))
        return out


def _make(name: str) -> type:
    tag, type_id, fields, derived, layout, needs_earth = SCHEMA[name]
    spec = [(f, float) for f in GEOMETRY_FIELDS + fields]
    cls = dataclasses.make_dataclass(
        name, spec, bases=(ZodiacalComponent,),
        namespace={"type_tag": tag, "type_id": type_id, "shape_layout": layout,
                   "needs_earth": needs_earth, "_derived_deg": derived,
                   "__doc__": f"{name} component (type '{tag}'); fields: {', '.join(GEOMETRY_FIELDS + fields)}."},
    )
    cls.__module__ = __name__
    return cls


Cloud = _make("Cloud")
Band = _make("Band")
Ring = _make("Ring")
Feature = _make("Feature")
Fan = _make("Fan")
Comet = _make("Comet")
Interstellar = _make("Interstellar")
NarrowBand = _make("NarrowBand")
BroadBand = _make("BroadBand")
RingRRM = _make("RingRRM")
FeatureRRM = _make("FeatureRRM")

COMPONENT_CLASSES = {SCHEMA[n][0]: globals()[n] for n in SCHEMA}
