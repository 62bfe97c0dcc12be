"""Host-side mirror of the reference interface: carried-over data, spectral unpacking, Model
constructor/validation semantics (reference tests/test_model.py, tests/test_evaluate.py)."""
import json
import os

import numpy as np
import pytest

import zodipy_b200 as zp
from helpers import GOLDEN_DIR, case_ids, golden_case
from zodipy_b200 import model_data as md
from zodipy_b200 import units as zu
from zodipy_b200.component import COMPONENT_CLASSES, SCHEMA, ComponentLabel
from zodipy_b200.zodiacal_light_model import ModelRegistry, model_registry

Q = zp.Quantity


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def test_carried_over_tables_equal_reference_dump():
    ref = json.load(open(os.path.join(GOLDEN_DIR, "reference_tables.json")))
    for set_name, mine in (("DIRBE", md.DIRBE_COMPONENTS), ("RRM", md.RRM_COMPONENTS)):
        assert list(mine) == list(ref["comps"][set_name])
        for label, (tag, fields) in mine.items():
            assert tag == ref["comps"][set_name][label]["type"]
            assert fields == ref["comps"][set_name][label]["fields"]
    assert list(ref["comps"]["PLANCK"]) == list(md.PLANCK_LABELS)
    src = ref["source"]
    for name in ("EMISSIVITY_DIRBE", "ALBEDO_DIRBE", "EMISSIVITY_PLANCK_13", "EMISSIVITY_PLANCK_15",
                 "EMISSIVITY_PLANCK_18", "EMISSIVITY_ODEGARD"):
        assert {k: list(v) for k, v in getattr(md, name).items()} == src[name]
    for name in ("C1_DIRBE", "C2_DIRBE", "C3_DIRBE", "SOLAR_IRRADIANCE_DIRBE", "CALIBRATION_RRM"):
        assert list(getattr(md, name)) == src[name]
    assert md.T_0_DIRBE == src["T_0_DIRBE"] and md.DELTA_DIRBE == src["DELTA_DIRBE"]
    assert md.T_0_RRM == src["T_0_RRM"] and md.DELTA_RRM == src["DELTA_RMM"]
    for name in ("SPECTRUM_DIRBE", "SPECTRUM_PLANCK", "SPECTRUM_IRAS"):
        assert list(getattr(md, name)[0]) == src[name]["value"]
        assert getattr(md, name)[1] == src[name]["unit"]
    assert {k: list(v) for k, v in md.COMPONENT_CUTOFFS.items()} == ref["cutoffs"]
    for name, info in ref["models"].items():
        m = model_registry.get_model(name)
        assert type(m).__name__ == info["class"]
        assert [k.value for k in m.comps] == info["comps"]
        assert list(zu.native_value(m.spectrum)) == info["spectrum"]


@pytest.mark.parametrize("case_id", [c for c in case_ids() if not golden_case(c)[0]["mutated"]])
def test_model_spec_matches_reference_unpack(case_id):
    """Model(...) prepares the same kernel inputs as the reference's __init__ path."""
    case, _ = golden_case(case_id)
    model = zp.Model(Q(case["x"], case["unit"]), weights=case["weights"], name=case["model"],
                     gauss_quad_degree=case["deg"], extrapolate=case["extrapolate"])
    mine, ref = model.spec, case["spec"]
    assert mine["kind"] == ref["kind"]
    for k in ("T_0", "delta", "C1", "C2", "C3", "solar_irradiance", "calibration"):
        assert (k in mine) == (k in ref)
        if k in ref:
            assert _rel(mine[k], ref[k]) <= 1e-14
    assert _rel(mine["table"], ref["table"]) <= 1e-14
    np.testing.assert_array_equal(mine["points"], ref["points"])
    np.testing.assert_array_equal(mine["weights"], ref["weights"])
    assert len(mine["comps"]) == len(ref["comps"])
    for cm, cr in zip(mine["comps"], ref["comps"]):
        assert (cm["label"], cm["type"], cm["cutoff"]) == (cr["label"], cr["type"], cr["cutoff"])
        for k in ("emissivity", "albedo", "T_0", "delta"):
            if k in cr:
                assert _rel(cm[k], cr[k]) <= 1e-14
        assert set(cm["params"]) == set(cr["params"])
        for k, v in cr["params"].items():
            assert _rel(cm["params"][k], v) == 0.0, (cm["label"], k)


def test_ghz_micron_parity():
    """reference tests/test_evaluate.py:33-45: 20 um == c/20um GHz."""
    m1 = zp.Model(Q(20.0, "um"))
    m2 = zp.Model(Q(zu.C_LIGHT / 20e-6 / 1e9, "GHz"))
    assert _rel(m1.spec["table"], m2.spec["table"]) <= 1e-13
    for a, b in zip(m1.spec["comps"], m2.spec["comps"]):
        assert _rel(a["emissivity"], b["emissivity"]) <= 1e-12


def test_x_input_errors():  # reference tests/test_model.py:11-31
    with pytest.raises(TypeError):
        zp.Model(20)
    with pytest.raises(TypeError):
        zp.Model(x=20)
    with pytest.raises(zu.UnitConversionError):
        Q(20, "s")  # not a wavelength / frequency unit
    with pytest.raises(ValueError):
        zp.Model(Q(0.5, "um"))  # outside the dirbe spectrum
    zp.Model(Q(0.5, "um"), extrapolate=True)
    with pytest.raises(ValueError):
        zp.Model(Q(30, "GHz"), name="planck18")


def test_weights_input_errors():  # reference tests/test_model.py:34-46
    with pytest.raises(ValueError):
        zp.Model(Q([20, 21, 22], "um"))
    with pytest.raises(ValueError):
        zp.Model(Q(20, "um"), weights=[1, 2, 3])
    with pytest.raises(ValueError):
        zp.Model(Q([20, 21, 22], "um"), weights=[1, 2])
    zp.Model(Q([20, 21, 22], "um"), weights=[1, 2, 1])


def test_unknown_model_and_registry():  # reference tests/test_model.py:75-91
    with pytest.raises(ValueError):
        zp.Model(Q(25, "um"), name="metamodel")
    with pytest.raises(ValueError):
        model_registry.get_model("metamodel")
    reg = ModelRegistry()
    base = model_registry.get_model("dirbe")
    reg.register_model("mine", base)
    assert reg.models == ["mine"]
    with pytest.raises(ValueError):
        reg.register_model("MINE", base)
    with pytest.raises(TypeError):
        reg.register_model("other", object())


def test_registry_models_are_not_mutated_through_model():
    """Model works on a private copy of the registered model."""
    m = zp.Model(Q(25, "um"))
    p = m.get_parameters()
    p["T_0"] += 250
    m.update_parameters(p)
    assert m.spec["T_0"] == 536
    assert model_registry.get_model("dirbe").T_0 == 286
    assert zp.Model(Q(25, "um")).spec["T_0"] == 286


def test_get_and_update_parameters_roundtrip():  # reference tests/test_model.py:54-72
    m = zp.Model(Q(25, "um"))
    p = m.get_parameters()
    assert set(p) >= {"comps", "spectrum", "T_0", "delta", "emissivities", "albedos"}
    assert p["comps"]["cloud"]["x_0"] == md.DIRBE_COMPONENTS["cloud"][1]["x_0"]
    assert list(p["comps"]["band1"]) == ["x_0", "y_0", "z_0", "i", "Omega", "n_0", "delta_zeta", "v",
                                         "p", "delta_r"]
    before = m.spec["comps"][0]["params"]["n_0"]
    p["comps"]["cloud"]["n_0"] = before * 2
    p["comps"]["band2"]["delta_zeta"] = 3.0
    m.update_parameters(p)
    assert m.spec["comps"][0]["params"]["n_0"] == before * 2
    assert m.spec["comps"][2]["params"]["delta_zeta_rad"] == pytest.approx(np.radians(3.0))
    assert m.get_parameters()["comps"]["band2"]["delta_zeta"] == 3.0
    # rrm model round-trips too (per-component T_0 / delta dicts keyed by label value)
    r = zp.Model(Q(25, "um"), name="rrm-experimental")
    pr = r.get_parameters()
    pr["T_0"]["fan"] = 300
    r.update_parameters(pr)
    assert r.spec["comps"][0]["T_0"] == 300


def test_component_schema_consistency():
    for name, (tag, type_id, fields, derived, layout, needs_earth) in SCHEMA.items():
        cls = COMPONENT_CLASSES[tag]
        assert cls.__name__ == name and cls.type_id == type_id
        assert len(layout) <= 8
        for f in layout:
            assert f in fields or f in derived
    assert len({v[1] for v in SCHEMA.values()}) == len(SCHEMA)
    assert ComponentLabel("inner_narrow_band") is ComponentLabel.INNER_NARROW_BAND


def test_units_shim():
    assert Q(25, "um").isscalar and not Q([1, 2], "um").isscalar
    assert Q(25, "um").to_value("micron") == 25
    assert zu.spectral_value(Q(857, "GHz"), "um") == pytest.approx(349.8161703617, rel=1e-12)
    with pytest.raises(zu.UnitConversionError):
        Q(1, "um").to_value("GHz")
    with pytest.raises(zu.UnitConversionError):
        Q(1, "parsec")
    assert zu.length_value(Q([1.0, 2.0], "AU"), "AU")[1] == 2.0


def test_evaluate_requires_skycoord_like():  # reference tests/test_evaluate.py:98-104
    m = zp.Model(Q(25, "um"))
    with pytest.raises(TypeError):
        m.evaluate([1.0, 2.0])

    class NoTime:
        obstime = None

    with pytest.raises(ValueError):
        m.evaluate(NoTime())


def test_lonlat_argument_packing():
    """The spherical-coordinate arguments reach the C struct unchanged (no GPU needed)."""
    import ctypes as C

    from zodipy_b200 import _cabi, engine

    lon, lat = np.linspace(0.0, 6.0, 7), np.linspace(-1.5, 1.5, 7)
    rot = np.arange(9.0).reshape(3, 3)
    ll = engine._LonLat(lon[::1], lat, rot)
    base = _cabi.EvalArgs()
    base.n = 7
    packed = ll.pack(base)
    assert packed.base.n == 7 and packed.has_rot == 1 and list(packed.rot) == list(range(9))
    assert packed.lon == ll.lon.ctypes.data and packed.lat == ll.lat.ctypes.data
    got = np.ctypeslib.as_array(C.cast(packed.lon, C.POINTER(C.c_double)), shape=(7,))
    np.testing.assert_array_equal(got, lon)
    assert engine._LonLat(lon, lat).pack(base).has_rot == 0
    # non-contiguous / non-float64 inputs are copied into contiguous float64
    ll2 = engine._LonLat(np.arange(14, dtype=np.float32)[::2], lat)
    assert ll2.lon.dtype == np.float64 and ll2.lon.flags.c_contiguous and ll2.n == 7
    with pytest.raises(ValueError):
        engine._LonLat(lon, lat[:-1])
    with pytest.raises(ValueError):
        engine._LonLat(lon, lat, np.eye(2))
