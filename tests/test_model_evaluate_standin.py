"""``Model.evaluate`` (the reference's public entry, zodipy/model.py:119-203) exercised end to end.

Astropy is not installed in this image, so a small duck-typed stand-in (tests/fake_astropy) plays
SkyCoord / Time / Quantity / get_body; with a real Astropy installed the stand-in is not used.  The
cases mirror the reference's tests/test_evaluate.py.  CPU variants replace the array seam by the
oracle (the glue above the seam is what is under test); GPU variants run the real kernels.
"""
import importlib
import json
import os
import sys

import numpy as np
import pytest

import zodi_oracle as oracle
from helpers import GOLDEN_DIR

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ap():
    """(astropy.units, astropy.time, astropy.coordinates), real if installed, else the stand-in."""
    try:
        import astropy  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(HERE, "fake_astropy"))
        importlib.invalidate_caches()
    from astropy import coordinates, time, units

    sys.modules.pop("zodipy_b200.astro", None)
    return units, time, coordinates


@pytest.fixture
def oracle_seam(monkeypatch):
    """Route the array seam to the oracle so the host glue can be tested without a GPU."""
    import zodipy_b200 as zp

    def seam(self, unit_vectors, obs_xyz, earth_xyz=None, *, return_comps=False, **kwargs):
        earth = obs_xyz if earth_xyz is None else earth_xyz
        em = oracle.evaluate(self.spec, np.asarray(unit_vectors), np.asarray(obs_xyz), np.asarray(earth))
        return em if return_comps else em.sum(axis=0)

    def lonlat_seam(self, lon, lat, obs_xyz=None, earth_xyz=None, *, frame_rotation=None, return_comps=False,
                    **kwargs):
        assert kwargs.get("ephemeris") is None
        lon, lat = np.asarray(lon), np.asarray(lat)
        u = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
        if frame_rotation is not None:
            u = np.asarray(frame_rotation) @ u
        return seam(self, u, obs_xyz, earth_xyz, return_comps=return_comps)

    monkeypatch.setattr(zp.Model, "evaluate_xyz", seam)
    monkeypatch.setattr(zp.Model, "evaluate_lonlat", lonlat_seam)


def _dirbe_table_check(ap, rel):
    import zodipy_b200 as zp

    units, time, coords = ap
    tab = json.load(open(os.path.join(GOLDEN_DIR, "dirbe_tabulated.json")))
    start = time.Time(tab["start_day"])
    for band, values in tab["emission"].items():
        model = zp.Model(float(band) * units.micron, name="dirbe")
        for day, lon, lat, expected in zip(tab["days"], tab["lon"], tab["lat"], values):
            obstime = start + time.TimeDelta(day - 1, format="jd")
            coord = coords.SkyCoord(lon, lat, unit=units.deg, frame=coords.BarycentricMeanEcliptic, obstime=obstime)
            emission = model.evaluate(coord)
            assert emission.shape == (1,)
            assert emission.value[0] == pytest.approx(expected, rel=rel), (band, day)


def test_dirbe_idl_table_through_public_api_cpu(ap, oracle_seam):
    """reference tests/test_evaluate.py:48-66 (stand-in ephemeris: 1 % holds)."""
    _dirbe_table_check(ap, 0.01)


@pytest.mark.gpu
def test_dirbe_idl_table_through_public_api_gpu(ap):
    _dirbe_table_check(ap, 0.01)


def test_validation_errors(ap, oracle_seam):
    """reference tests/test_evaluate.py:98-133."""
    import zodipy_b200 as zp

    units, time, coords = ap
    model = zp.Model(25 * units.micron)
    t = time.Time("2021-01-01T00:00:00")
    with pytest.raises(TypeError):
        model.evaluate(20)
    with pytest.raises(ValueError):
        model.evaluate(coords.SkyCoord(20, 30, unit=units.deg))  # no obstime
    sc = coords.SkyCoord(20, 30, unit=units.deg, obstime=t)
    with pytest.raises(TypeError):
        model.evaluate(sc, obspos=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        model.evaluate(sc, obspos="not-a-body")
    with pytest.raises(units.UnitConversionError):
        model.evaluate(sc, obspos=[0.1, 0.2, 1.0] * units.s)
    with pytest.raises(ValueError):  # (3, 2) positions for a single obstime
        model.evaluate(sc, obspos=np.ones((3, 2)) * units.AU)
    with pytest.raises(ValueError):  # more obstimes than coordinates
        model.evaluate(coords.SkyCoord(20, 30, unit=units.deg, obstime=time.Time([59000.0, 59001.0], format="mjd")))


def test_shapes_and_return_comps(ap, oracle_seam):
    """reference tests/test_evaluate.py:135-212."""
    import zodipy_b200 as zp

    units, time, coords = ap
    t = time.Time("2021-01-01T00:00:00")
    model = zp.Model(25 * units.micron)
    for frame in ("galactic", "icrs", coords.BarycentricMeanEcliptic):
        scalar = coords.SkyCoord(20, 30, unit=units.deg, obstime=t, frame=frame)
        assert model.evaluate(scalar).shape == (1,)
        assert model.evaluate(scalar, return_comps=True).shape == (6, 1)
    many = coords.SkyCoord([10, 10.1, 10.2], [90, 89, 77], unit=units.deg, obstime=t, frame="galactic")
    total = model.evaluate(many, nprocesses=4)  # nprocesses accepted, result identical
    comps = model.evaluate(many, return_comps=True)
    assert total.shape == (3,) and comps.shape == (6, 3)
    assert str(total.unit) == "MJy / sr"
    np.testing.assert_array_equal(np.asarray(comps.sum(axis=0)), np.asarray(total))
    np.testing.assert_array_equal(np.asarray(model.evaluate(many, nprocesses=1)), np.asarray(total))
    # explicit observer position, Mars, SEMB-L2
    for obspos in ([0.87, -0.53, 0.001] * units.AU, "mars", "semb-l2"):
        assert model.evaluate(many, obspos=obspos).shape == (3,)
    far = model.evaluate(many, obspos="mars")
    assert np.all(np.asarray(far) < np.asarray(total))  # fainter from 1.5 AU (docs/usage.md:199)


def test_sky_rotation_is_read_off_the_frame_transformation(ap, oracle_seam):
    """Model(sky_rotation="device") ships SkyCoord angles + one 3x3 matrix (taken from the frame
    transformation itself) instead of host-transformed unit vectors; same results as "host"."""
    import zodipy_b200 as zp
    from zodipy_b200 import astro

    units, time, coords = ap
    t = time.Time("2021-01-01T00:00:00")
    rng = np.random.default_rng(3)
    lon, lat = rng.uniform(0, 360, 50), rng.uniform(-90, 90, 50)
    for frame in ("galactic", "icrs", coords.BarycentricMeanEcliptic):
        sc = coords.SkyCoord(lon, lat, unit=units.deg, obstime=t, frame=frame)
        got = astro.sky_lonlat_rotation(sc)
        assert got is not None
        lon_r, lat_r, rot = got
        np.testing.assert_allclose(lon_r, np.radians(lon), rtol=0, atol=1e-15)
        u = rot @ np.array([np.cos(lat_r) * np.cos(lon_r), np.cos(lat_r) * np.sin(lon_r), np.sin(lat_r)])
        np.testing.assert_allclose(u, astro.sky_unit_vectors(sc), rtol=0, atol=1e-15)
        dev = zp.Model(25 * units.micron, sky_rotation="device").evaluate(sc, return_comps=True)
        host = zp.Model(25 * units.micron, sky_rotation="host").evaluate(sc, return_comps=True)
        np.testing.assert_allclose(np.asarray(dev), np.asarray(host), rtol=1e-12)
    # a coordinate that is not a pure direction in a known frame -> host transformation
    moved = coords.SkyCoord(lon, lat, unit=units.deg, obstime=t, frame="galactic").transform_to("icrs")
    assert astro.sky_lonlat_rotation(moved) is None or isinstance(moved.data, coords.UnitSphericalRepresentation)
    assert zp.Model(25 * units.micron).evaluate(moved).shape == (50,)
    with pytest.raises(ValueError):
        zp.Model(25 * units.micron, sky_rotation="gpu")


def test_sky_rotation_with_per_sample_frame_attributes(ap, oracle_seam):
    """Time-ordered SkyCoords whose FRAME carries the N-element obstime (Astropy frames such as
    HeliocentricMeanEcliptic do): the probe directions are attached to a copy of the frame with scalar
    attributes, so the device-rotation path is taken; a frame whose orientation really depends on the
    per-sample attribute is detected and falls back to the host transformation with a warning."""
    import zodipy_b200 as zp
    from zodipy_b200 import astro

    units, time, coords = ap
    if not hasattr(coords, "ObstimeEcliptic"):
        pytest.skip("stand-in frames (real Astropy is covered by tools/verify_with_astropy.py)")
    n = 40
    times = time.Time(59215.0 + np.linspace(0.0, 30.0, n), format="mjd")
    rng = np.random.default_rng(8)
    lon, lat = rng.uniform(0, 360, n), rng.uniform(-90, 90, n)
    sc = coords.SkyCoord(lon, lat, unit=units.deg, frame=coords.ObstimeEcliptic(obstime=times))
    assert sc.frame.obstime.size == n and sc.obstime.size == n
    got = astro.sky_lonlat_rotation(sc)
    assert got is not None, "per-sample frame attributes must not disable the device rotation"
    np.testing.assert_allclose(got[2], np.eye(3), rtol=0, atol=1e-15)
    dev = zp.Model(25 * units.micron, sky_rotation="device").evaluate(sc)
    host = zp.Model(25 * units.micron, sky_rotation="host").evaluate(sc)
    np.testing.assert_allclose(np.asarray(dev), np.asarray(host), rtol=1e-12)
    turning = coords.SkyCoord(lon, lat, unit=units.deg, frame=coords.TimeDependentFrame(obstime=times))
    with pytest.warns(RuntimeWarning, match="not a fixed rotation"):
        assert astro.sky_lonlat_rotation(turning) is None
    with pytest.warns(RuntimeWarning):
        fallback = zp.Model(25 * units.micron, sky_rotation="device").evaluate(turning)
    np.testing.assert_allclose(np.asarray(fallback),
                               np.asarray(zp.Model(25 * units.micron, sky_rotation="host").evaluate(turning)), rtol=1e-12)


def test_time_ordered_inputs(ap, oracle_seam):
    import zodipy_b200 as zp

    units, time, coords = ap
    n = 40
    times = time.Time(59215.0 + np.linspace(0, 30, n), format="mjd")
    rng = np.random.default_rng(1)
    sc = coords.SkyCoord(rng.uniform(0, 360, n), rng.uniform(-90, 90, n), unit=units.deg, obstime=times,
                         frame="galactic")
    model = zp.Model(25 * units.micron)
    base = model.evaluate(sc)
    assert base.shape == (n,)
    # per-sample explicit positions == the same positions derived from the ephemeris
    from zodipy_b200 import astro

    knots = astro.arrange_obstimes(times[0].mjd, times[-1].mjd)
    earth = astro.interp_bodypos("earth", times.mjd, knots, "builtin")
    same = model.evaluate(sc, obspos=earth * units.AU)
    np.testing.assert_allclose(np.asarray(same), np.asarray(base), rtol=1e-13)
    with pytest.raises(ValueError):
        model.evaluate(sc, obspos=earth[:, :-1] * units.AU)
    l2 = model.evaluate(sc, obspos="semb-l2")
    assert not np.allclose(np.asarray(l2), np.asarray(base))
    assert model.evaluate(sc, obspos="mars", return_comps=True).shape == (6, n)


@pytest.mark.gpu
@pytest.mark.parametrize("obspos", ["earth", "semb-l2", "mars"])
def test_tod_device_ephemeris_equals_host_path_gpu(ap, obspos):
    """Model(tod_ephemeris="device") == the reference-style host interpolation, real kernels."""
    import zodipy_b200 as zp

    units, time, coords = ap
    n = 5000
    times = time.Time(59215.0 + np.linspace(0, 20, n), format="mjd")
    rng = np.random.default_rng(2)
    sc = coords.SkyCoord(rng.uniform(0, 360, n), rng.uniform(-90, 90, n), unit=units.deg, obstime=times,
                         frame="galactic")
    host = zp.Model(25 * units.micron).evaluate(sc, obspos=obspos, return_comps=True)
    dev = zp.Model(25 * units.micron, tod_ephemeris="device").evaluate(sc, obspos=obspos, return_comps=True)
    assert dev.shape == (6, n)
    np.testing.assert_allclose(np.asarray(dev), np.asarray(host), rtol=1e-11, atol=1e-30)
    with pytest.raises(ValueError):
        zp.Model(25 * units.micron, tod_ephemeris="device").evaluate(sc, obspos="not-a-body")


@pytest.mark.gpu
@pytest.mark.parametrize("tod", [False, True])
def test_sky_rotation_device_equals_host_gpu(ap, tod):
    """SkyCoord angles rotated in the kernel prologue == Astropy-transformed unit vectors, real kernels."""
    import zodipy_b200 as zp

    units, time, coords = ap
    n = 4000
    rng = np.random.default_rng(8)
    times = time.Time(59215.0 + np.linspace(0, 20, n), format="mjd") if tod else time.Time("2021-01-01T00:00:00")
    for frame in ("galactic", "icrs", coords.BarycentricMeanEcliptic):
        sc = coords.SkyCoord(rng.uniform(0, 360, n), rng.uniform(-90, 90, n), unit=units.deg, obstime=times,
                             frame=frame)
        for kw in ({}, {"tod_ephemeris": "device"}):
            host = zp.Model(25 * units.micron, sky_rotation="host", **kw).evaluate(sc, obspos="semb-l2")
            dev = zp.Model(25 * units.micron, sky_rotation="device", **kw).evaluate(sc, obspos="semb-l2")
            assert dev.shape == (n,)
            np.testing.assert_allclose(np.asarray(dev), np.asarray(host), rtol=1e-11)


@pytest.mark.gpu
def test_ghz_micron_parity_gpu(ap):
    """reference tests/test_evaluate.py:33-45."""
    import zodipy_b200 as zp

    units, time, coords = ap
    t = time.Time("2021-01-01T00:00:00")
    sc = coords.SkyCoord([10, 10.1, 10.2], [90, 89, 77], unit=units.deg, obstime=t, frame="galactic")
    a = zp.Model(20.0 * units.micron).evaluate(sc)
    b = zp.Model((299792458.0 / 20e-6 / 1e9) * units.GHz).evaluate(sc)
    assert [round(v, 12) for v in np.asarray(a)] == [round(v, 12) for v in np.asarray(b)]
