"""Loader for the UNMODIFIED reference hot-path modules (test infrastructure only).

This file is test infrastructure: it is used ONLY by ``oracle/make_golden.py`` (run in the
build container, where ``/root/reference`` exists) to generate the committed fixtures under
``tests/golden/`` and to validate the NumPy restatement in ``oracle/zodi_oracle.py``.
Nothing in the product (``zodipy_b200/``), in ``bench.py`` or in the ``-m gpu`` tests imports it,
and it is never executed on the GPU box (``/root/reference`` does not exist there).

``import zodipy`` fails in this image because Astropy is not installed
(``zodipy/__init__.py:1`` -> ``zodipy/model.py:10-11``).  The six array-only hot-path files do not
need Astropy, so they are imported unchanged by registering an empty package object whose
``__path__`` points at the reference and pre-seeding four stub sibling modules
(SURVEY.md Appendix D):

* ``zodipy.blackbody``  -> only ``get_dust_grain_temperature`` (``zodipy/blackbody.py:30``)
* ``zodipy.bodies``, ``zodipy.model_registry``, ``zodipy.zodiacal_light_model`` -> empty stubs
  (they are imported by ``zodipy/number_density.py:12,28,29`` only for ``grid_number_density``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ZODIPY_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "zodipy"))


def load():
    """Return a namespace with the reference's hot-path modules (imported unchanged)."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if "zodipy" in sys.modules and not getattr(sys.modules["zodipy"], "_oracle_stub", False):
        raise RuntimeError("a real `zodipy` package is already imported")

    pkg = types.ModuleType("zodipy")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "zodipy")]
    pkg._oracle_stub = True
    sys.modules["zodipy"] = pkg

    bb = types.ModuleType("zodipy.blackbody")
    bb.get_dust_grain_temperature = lambda R, T_0, delta: T_0 * R**-delta  # blackbody.py:30
    sys.modules["zodipy.blackbody"] = bb

    bodies = types.ModuleType("zodipy.bodies")
    bodies.get_earthpos_inst = None
    sys.modules["zodipy.bodies"] = bodies

    reg = types.ModuleType("zodipy.model_registry")
    reg.model_registry = None
    sys.modules["zodipy.model_registry"] = reg

    zlm = types.ModuleType("zodipy.zodiacal_light_model")
    zlm.ZodiacalLightModel = object
    sys.modules["zodipy.zodiacal_light_model"] = zlm

    ns = types.SimpleNamespace()
    for name in ("component", "component_params", "scattering", "line_of_sight",
                 "brightness", "number_density"):
        setattr(ns, name, importlib.import_module(f"zodipy.{name}"))
    return ns
