"""Pins the CPU oracle (oracle/zodi_oracle.py) to outputs of the reference's own modules.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which runs the
UNMODIFIED reference hot-path modules on fixed inputs.
"""
import json
import os

import numpy as np
import pytest

import zodi_oracle as oracle
from helpers import GOLDEN_DIR, case_ids, golden_case, max_rel_comps, max_rel_total


@pytest.mark.parametrize("case_id", case_ids())
def test_oracle_matches_reference_output(case_id):
    case, a = golden_case(case_id)
    em = oracle.evaluate(case["spec"], a["u"], a["obs"], a["earth"])
    assert em.shape == a["emission"].shape
    assert max_rel_total(em, a["emission"]) <= 1e-12
    assert max_rel_comps(em, a["emission"]) <= 1e-12


@pytest.mark.parametrize("case_id", ["g1_dirbe25_fix", "rrm_25um_obs1p63", "dirbe_25um_tod_straddle",
                                     "dirbe_25um_obs0p74"])
def test_oracle_range_matches_reference(case_id):
    case, a = golden_case(case_id)
    start, stop = oracle.los_range(case["spec"], a["u"], a["obs"])
    n = a["u"].shape[1]
    np.testing.assert_allclose(np.array([np.broadcast_to(s, n) for s in start]), a["start"], rtol=1e-14)
    np.testing.assert_allclose(np.array([np.broadcast_to(s, n) for s in stop]), a["stop"], rtol=1e-14)


def test_survey_appendix_d_goldens():
    """Known answers quoted in SURVEY.md Appendix D (independent survey-time run)."""
    _, a = golden_case("g1_dirbe25_fix")
    np.testing.assert_allclose(
        a["emission"].sum(axis=0),
        [66.98621366983629, 17.89508455836402, 14.533643688599547, 393.8965068277863], rtol=1e-12)
    np.testing.assert_allclose(
        a["emission"][:, 0],
        [61.11018696340161, 0.28969398892761494, 3.4057290180826025, 0.07448231171041658,
         0.8489545873465021, 1.257166800367556], rtol=1e-12)
    np.testing.assert_allclose(
        a["stop"][0], [5.289296526881172, 4.467931878080205, 5.134261550536542, 6.105268343284958],
        rtol=1e-13)
    _, a = golden_case("g2_dirbe1p25_fix")
    np.testing.assert_allclose(
        a["emission"].sum(axis=0),
        [0.46894217513427866, 0.13695230983101903, 0.10836675063421672, 7.982005929343584], rtol=1e-12)
    _, a = golden_case("g3_planck18_857_fix")
    np.testing.assert_allclose(
        a["emission"].sum(axis=0),
        [0.4221879039421093, 0.10024945486894771, 0.06997615992841576, 1.4387064776316687], rtol=1e-12)
    _, a = golden_case("g4_rrm25_fix")
    np.testing.assert_allclose(
        a["emission"].sum(axis=0),
        [25.959312780951233, 7.232840630735481, 6.031681558626123, 192.28793347536978], rtol=1e-12)


def test_blackbody_table_known_values():
    """Table check values of SURVEY Appendix D (25 um)."""
    tab = oracle.blackbody_table(oracle.C_LIGHT / 25e-6)
    assert tab.shape == (2, 100)
    assert tab[0, 0] == 40.0 and tab[0, -1] == 550.0
    np.testing.assert_allclose(tab[1, 0], 1434.6909578113907, rtol=1e-12)
    np.testing.assert_allclose(tab[1, 1], 7407.620931981804, rtol=1e-12)
    np.testing.assert_allclose(tab[1, -1], 1376389131.2942388, rtol=1e-12)


def test_interp_spectral_param_matches_scipy():
    """The interp1d restatement against SciPy itself (linear, nearest incl. ties, extrapolation)."""
    from scipy import interpolate

    rng = np.random.default_rng(3)
    knots = np.array([1.25, 2.2, 3.5, 4.9, 12, 25, 60, 100, 140, 240.0])
    vals = rng.normal(size=knots.size)
    xq = np.concatenate([knots, 0.5 * (knots[1:] + knots[:-1]), rng.uniform(1.25, 240, 200)])
    for nearest in (False, True):
        f = interpolate.interp1d(knots, vals, kind="nearest" if nearest else "linear")
        np.testing.assert_allclose(
            oracle.interp_spectral_param(xq, None, knots, vals, use_nearest=nearest), f(xq), rtol=1e-14)
        f = interpolate.interp1d(knots, vals, kind="nearest" if nearest else "linear",
                                 bounds_error=False, fill_value="extrapolate")
        xe = np.array([0.5, 1.0, 300.0, 1000.0])
        np.testing.assert_allclose(
            oracle.interp_spectral_param(xe, None, knots, vals, use_nearest=nearest, bounds_error=False),
            f(xe), rtol=1e-14)
    # descending spectrum is flipped first (unpack_model.py:151-153)
    np.testing.assert_allclose(
        oracle.interp_spectral_param(3.0, None, knots[::-1], vals[::-1]),
        oracle.interp_spectral_param(3.0, None, knots, vals))
    with pytest.raises(ValueError):
        oracle.interp_spectral_param(0.5, None, knots, vals, bounds_error=True)


def _earth_analytic(jd):
    """Low-precision heliocentric mean-ecliptic (J2000) Earth position (Meeus ch. 25) [AU]."""
    T = (jd - 2451545.0) / 36525.0
    L0 = 280.46646 + 36000.76983 * T + 0.0003032 * T * T
    M = np.radians(357.52911 + 35999.05029 * T - 0.0001537 * T * T)
    e = 0.016708634 - 0.000042037 * T
    Cc = ((1.914602 - 0.004817 * T) * np.sin(M) + (0.019993 - 0.000101 * T) * np.sin(2 * M)
          + 0.000289 * np.sin(3 * M))
    sun_lon = L0 + Cc
    nu = M + np.radians(Cc)
    R = 1.000001018 * (1 - e * e) / (1 + e * np.cos(nu))
    lon = np.radians(sun_lon + 180.0 - 1.396971 * T)  # Earth = Sun + 180 deg; precess to J2000
    return np.array([[R * np.cos(lon)], [R * np.sin(lon)], [0.0]])


def test_dirbe_idl_table_within_reference_tolerance():
    """The reference's own known-answer test (tests/test_evaluate.py:48-66): DIRBE IDL software
    values within 1 %.  Earth from an analytic ephemeris (Astropy is absent here)."""
    tab = json.load(open(os.path.join(GOLDEN_DIR, "dirbe_tabulated.json")))
    jd0 = 2447892.5  # 1990-01-01T00:00
    specs = {c["spec"]["name"] + str(c["x"]): c["spec"] for c in map(lambda i: golden_case(i)[0], case_ids())
             if c["model"] == "dirbe" and c["weights"] is None and not c["mutated"] and c["deg"] == 50}
    for band, values in tab["emission"].items():
        spec = specs["dirbe" + str(float(band))]
        for day, lon, lat, expected in zip(tab["days"], tab["lon"], tab["lat"], values):
            earth = _earth_analytic(jd0 + day - 1)
            lo, la = np.radians(lon), np.radians(lat)
            u = np.array([[np.cos(la) * np.cos(lo)], [np.cos(la) * np.sin(lo)], [np.sin(la)]])
            total = oracle.evaluate(spec, u, earth, earth).sum(axis=0)[0]
            assert total == pytest.approx(expected, rel=0.01), (band, day)


def test_oracle_parallel_driver_bitwise_equals_serial():
    """The nprocesses path (model.py:182-198) must equal the serial one bit for bit
    (reference: tests/test_evaluate.py:215-262)."""
    case, a = golden_case("dirbe_25um_tod")
    serial = oracle.evaluate(case["spec"], a["u"], a["obs"], a["earth"])
    par = oracle.evaluate_parallel(case["spec"], a["u"], a["obs"], a["earth"], nprocesses=3)
    np.testing.assert_array_equal(serial, par)
