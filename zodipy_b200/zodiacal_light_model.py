"""Interplanetary-dust model descriptions and the model registry (host configuration).

Mirrors the public surface of ``zodipy/zodiacal_light_model.py:20-142`` and
``zodipy/model_registry.py:4-73``: ``Kelsall`` / ``RRM`` model classes with ``to_dict`` /
``ncomps`` / ``is_valid_at`` and a ``model_registry`` holding the six shipped models.  The
``brightness_at_step_callable`` property of the reference (which selects the Python source
function) is replaced by ``kind``: the device kernel takes the Kelsall/RRM distinction as data.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Mapping, Sequence

import numpy as np

from . import model_data as md
from . import units as zu
from .component import COMPONENT_CLASSES, ComponentLabel, ZodiacalComponent


@dataclass(repr=False)
class ZodiacalLightModel:
    """Base class: component mapping + the spectrum over which the model is tabulated."""

    comps: Mapping[ComponentLabel, ZodiacalComponent]
    spectrum: zu.Quantity

    kind = ""

    def to_dict(self) -> dict:
        """Nested plain-dict form (``zodiacal_light_model.py:38-55``)."""
        out: dict = {}
        for key, value in vars(self).items():
            if key == "comps":
                out[key] = {label.value: comp.init_fields() for label, comp in value.items()}
            elif isinstance(value, dict):
                out[key] = {k.value: v for k, v in value.items()}
            else:
                out[key] = value
        return out

    @property
    def ncomps(self) -> int:
        return len(self.comps)

    def is_valid_at(self, x) -> bool:
        """True if every requested wavelength/frequency lies inside the tabulated spectrum."""
        try:
            xv = np.asarray(zu.spectral_value(x, zu.unit_name(self.spectrum)))
        except zu.UnitConversionError as err:
            raise zu.UnitConversionError("Input 'x' must have units convertible to Hz or m.") from err
        sv = zu.native_value(self.spectrum)
        return bool(np.all((sv.min() <= xv) & (xv <= sv.max())))


@dataclass(repr=False)
class Kelsall(ZodiacalLightModel):
    """Kelsall et al. (1998) type model (dirbe, planck13/15/18, odegard)."""

    T_0: float = 0.0
    delta: float = 0.0
    emissivities: Mapping[ComponentLabel, Sequence[float]] = field(default_factory=dict)
    albedos: Mapping[ComponentLabel, Sequence[float]] | None = None
    solar_irradiance: Sequence[float] | None = None
    C1: Sequence[float] | None = None
    C2: Sequence[float] | None = None
    C3: Sequence[float] | None = None

    kind = "kelsall"


@dataclass(repr=False)
class RRM(ZodiacalLightModel):
    """Rowan-Robinson and May (2013) type model (rrm-experimental)."""

    T_0: Mapping[ComponentLabel, float] = field(default_factory=dict)
    delta: Mapping[ComponentLabel, float] = field(default_factory=dict)
    calibration: Sequence[float] = ()

    kind = "rrm"


class ModelRegistry:
    """Name -> model container (``zodiacal_light_model.py:105-142``)."""

    def __init__(self) -> None:
        self._registry: dict[str, ZodiacalLightModel] = {}

    @property
    def models(self) -> list[str]:
        return list(self._registry)

    def register_model(self, name: str, model: ZodiacalLightModel) -> None:
        key = name.lower()
        if key in self._registry:
            raise ValueError(f"a model by the name {key!s} is already registered.")
        if not isinstance(model, ZodiacalLightModel):
            raise TypeError("model must be an instance of ZodiacalLightModel.")
        self._registry[key] = model

    def get_model(self, name: str) -> ZodiacalLightModel:
        key = name.lower()
        if key not in self._registry:
            raise ValueError(
                f"{key!r} is not a registered Interplanetary Dust model. "
                f"Avaliable models are: {', '.join(self._registry)}."
            )
        return self._registry[key]


def _components(table: dict, labels=None) -> dict:
    out = {}
    for label, (tag, fields) in table.items():
        if labels is None or label in labels:
            out[ComponentLabel(label)] = COMPONENT_CLASSES[tag](**fields)
    return out


def _by_label(table: dict) -> dict:
    return {ComponentLabel(k): v for k, v in table.items()}


def _default_registry() -> ModelRegistry:
    reg = ModelRegistry()
    dirbe_spec = zu.Quantity(*md.SPECTRUM_DIRBE)
    planck_spec = zu.Quantity(*md.SPECTRUM_PLANCK)
    reg.register_model("dirbe", Kelsall(
        comps=_components(md.DIRBE_COMPONENTS), spectrum=dirbe_spec,
        emissivities=_by_label(md.EMISSIVITY_DIRBE), albedos=_by_label(md.ALBEDO_DIRBE),
        solar_irradiance=md.SOLAR_IRRADIANCE_DIRBE, C1=md.C1_DIRBE, C2=md.C2_DIRBE, C3=md.C3_DIRBE,
        T_0=md.T_0_DIRBE, delta=md.DELTA_DIRBE))
    reg.register_model("planck13", Kelsall(
        comps=_components(md.DIRBE_COMPONENTS), spectrum=planck_spec,
        emissivities=_by_label(md.EMISSIVITY_PLANCK_13), T_0=md.T_0_DIRBE, delta=md.DELTA_DIRBE))
    for name, emis in (("planck15", md.EMISSIVITY_PLANCK_15), ("planck18", md.EMISSIVITY_PLANCK_18),
                       ("odegard", md.EMISSIVITY_ODEGARD)):
        reg.register_model(name, Kelsall(
            comps=_components(md.DIRBE_COMPONENTS, md.PLANCK_LABELS), spectrum=planck_spec,
            emissivities=_by_label(emis), T_0=md.T_0_DIRBE, delta=md.DELTA_DIRBE))
    reg.register_model("rrm-experimental", RRM(
        comps=_components(md.RRM_COMPONENTS), spectrum=zu.Quantity(*md.SPECTRUM_IRAS),
        calibration=md.CALIBRATION_RRM, T_0=_by_label(md.T_0_RRM), delta=_by_label(md.DELTA_RRM)))
    return reg


model_registry = _default_registry()


def clone_model(model: ZodiacalLightModel) -> ZodiacalLightModel:
    """Private copy for a ``Model`` instance (the reference hands out the shared object, which
    lets one user's in-place edit leak into every later ``Model``)."""
    return copy.deepcopy(model)
