"""Astropy-side host preparation for ``Model.evaluate`` (imported lazily; needs Astropy).

Ephemerides stay on the Python host by design (BASELINE north star).  This follows
``zodipy/bodies.py:16-99`` and ``zodipy/model.py:212-251``: Earth / observer heliocentric
mean-ecliptic positions (hourly knots + cubic spline for time-ordered data, the SEMB-L2
approximation) and the sky coordinates as ``BarycentricMeanEcliptic`` unit vectors - either
transformed by Astropy on the host (``sky_unit_vectors``) or, when the frame differs from the mean
ecliptic by a fixed rotation, handed to the device as angles + that rotation
(``sky_lonlat_rotation``).
"""
from __future__ import annotations

import numpy as np

try:
    import astropy.coordinates as coords
    from astropy import time, units
except ImportError as err:  # pragma: no cover - astropy is absent from the build image
    raise ImportError(
        "Model.evaluate(SkyCoord) needs astropy for ephemerides and frame rotation; "
        "install astropy or call Model.evaluate_xyz() with ecliptic unit vectors.") from err

MEAN_DIST_TO_L2 = 0.009896235034000056  # AU, zodipy/bodies.py:13


def arrange_obstimes(t0: float, t1: float):
    """Hourly knots spanning the observation (``bodies.py:16-19``)."""
    dt = 1.0 / 24.0
    return time.Time(np.arange(t0, t1 + dt, dt), format="mjd")


def _body_xyz(body, obstime, ephemeris):
    return (coords.get_body(body, obstime, ephemeris=ephemeris)
            .transform_to(coords.HeliocentricMeanEcliptic).cartesian.xyz.to_value(units.AU))


def interp_bodypos(body, obstimes_mjd, interp_obstimes, ephemeris):
    """Cubic-spline interpolation of hourly positions (``bodies.py:22-35``)."""
    from scipy import interpolate

    pos = _body_xyz(body, interp_obstimes, ephemeris)
    return interpolate.CubicSpline(interp_obstimes.mjd, pos, axis=-1)(obstimes_mjd)


def semb_l2(earthpos):
    """SEMB-L2 approximation incl. the reference's un-axised norm (``bodies.py:38-50``, Q5)."""
    dist = np.linalg.norm(earthpos)
    return earthpos / dist * (dist + MEAN_DIST_TO_L2)


_PROBE_LON = np.array([0.0, 0.5 * np.pi, 0.0, 0.7, 4.1])
_PROBE_LAT = np.array([0.0, 0.0, 0.5 * np.pi, -0.4, 1.1])


def _scalar_frame(frame, pick):
    """Data-less copy of ``frame`` in which every array-valued frame attribute (e.g. a per-sample
    ``obstime``) is replaced by its element ``pick`` (0 or -1), so that a few probe directions can be
    attached to it.  Returns (frame, any attribute was an array)."""
    names = getattr(frame, "frame_attributes", None)
    if names is None:  # old Astropy spelling
        names = frame.get_frame_attr_names()
    attrs, varying = {}, False
    for name in names:
        value = getattr(frame, name)
        if getattr(value, "shape", ()) not in ((), None) and getattr(value, "size", 1) > 1:
            value, varying = value[pick], True
        attrs[name] = value
    return type(frame)(**attrs), varying


def _probe_rotation(frame):
    """(R, ok): the 3x3 matrix Astropy's transformation frame -> BarycentricMeanEcliptic applies to the
    frame's axes, and whether two further directions confirm that the transformation IS that rotation."""
    probe = coords.SkyCoord(_PROBE_LON * units.rad, _PROBE_LAT * units.rad, frame=frame)
    xyz = np.asarray(probe.transform_to(coords.BarycentricMeanEcliptic).cartesian.xyz.value, dtype=np.float64)
    rot = np.ascontiguousarray(xyz[:, :3])  # columns = images of the frame's x, y, z axes
    cl = np.cos(_PROBE_LAT[3:])
    src = np.array([cl * np.cos(_PROBE_LON[3:]), cl * np.sin(_PROBE_LON[3:]), np.sin(_PROBE_LAT[3:])])
    ok = (np.allclose(rot.T @ rot, np.eye(3), rtol=0, atol=1e-12)
          and np.allclose(rot @ src, xyz[:, 3:], rtol=0, atol=1e-12))
    return rot, bool(ok)


def sky_lonlat_rotation(skycoord):
    """``(lon, lat, R)`` with the mean-ecliptic unit vectors equal to
    ``R @ (cos lat cos lon, cos lat sin lon, sin lat)``, or ``None``.

    ``skycoord.transform_to(BarycentricMeanEcliptic).cartesian.xyz`` (``model.py:247-251``) is, for
    direction-only coordinates in ICRS / Galactic / FK5 / mean-ecliptic frames, a fixed rotation of
    the sphere.  R is read off Astropy itself: probe directions are attached to a data-less copy of the
    coordinate's frame (array-valued frame attributes such as the per-sample ``obstime`` of time-ordered
    data reduced to their first - and, for comparison, last - element), the images of the three axes
    give R, two more directions verify that the transformation really is that rotation, and finally a
    handful of the ACTUAL coordinates are transformed by Astropy and compared with ``R @ vector``.
    Frames that fail any of these (aberration, time-dependent orientation, distances) make the caller
    fall back to transforming every coordinate on the host, with a warning saying why.  The angles then
    go to the device as they are (16 B per line of sight) and the kernel prologue does the trigonometry.
    """
    import warnings

    def give_up(why):
        warnings.warn(f"sky_rotation='device' is not applicable ({why}); coordinates are transformed on the "
                      "host as in the reference", RuntimeWarning, stacklevel=3)
        return None

    data = skycoord.data
    if not isinstance(data, coords.UnitSphericalRepresentation):
        return None  # distances / cartesian data: not a pure direction, the reference path handles it
    lon = np.ascontiguousarray(np.atleast_1d(data.lon.to_value(units.rad)), dtype=np.float64).reshape(-1)
    lat = np.ascontiguousarray(np.atleast_1d(data.lat.to_value(units.rad)), dtype=np.float64).reshape(-1)
    try:
        frame0, varying = _scalar_frame(skycoord.frame, 0)
        rot, ok = _probe_rotation(frame0)
        if ok and varying:  # per-sample frame attributes: the rotation must not depend on them
            rot_last, ok_last = _probe_rotation(_scalar_frame(skycoord.frame, -1)[0])
            ok = ok_last and np.allclose(rot, rot_last, rtol=0, atol=1e-12)
    except (TypeError, ValueError, AttributeError, KeyError, IndexError, coords.ConvertError) as error:
        return give_up(f"{type(error).__name__}: {error}")
    if not ok:
        return give_up("the frame's transformation to BarycentricMeanEcliptic is not a fixed rotation")
    # last line of defence: Astropy's own transformation of a few of the actual coordinates
    n = lon.size
    if n > 0 and not getattr(skycoord, "isscalar", False):
        idx = np.unique(np.linspace(0, n - 1, min(n, 8)).astype(np.int64))
        try:
            want = np.asarray(skycoord[idx].transform_to(coords.BarycentricMeanEcliptic).cartesian.xyz.value,
                              dtype=np.float64).reshape(3, -1)
        except (TypeError, ValueError, AttributeError, KeyError, IndexError, coords.ConvertError) as error:
            return give_up(f"{type(error).__name__}: {error}")
        cl = np.cos(lat[idx])
        have = rot @ np.array([cl * np.cos(lon[idx]), cl * np.sin(lon[idx]), np.sin(lat[idx])])
        if not np.allclose(have, want, rtol=0, atol=1e-11):
            return give_up("Astropy's transformation of the coordinates differs from the probed rotation")
    return lon, lat, rot


def sky_unit_vectors(skycoord):
    """Mean-ecliptic unit vectors (3, N) computed by Astropy on the host (``model.py:247-251``)."""
    ecl = skycoord.transform_to(coords.BarycentricMeanEcliptic)
    u_xyz = ecl.cartesian.xyz.value
    if ecl.isscalar:
        u_xyz = u_xyz[:, np.newaxis]
    return np.ascontiguousarray(u_xyz)


def prepare_arrays(skycoord, obspos, obspos_isstr, interp_obstimes, ephemeris, with_directions=True):
    """(earth_xyz, obs_xyz, unit_vectors) as the array seam expects (``model.py:212-251``);
    ``unit_vectors`` is ``None`` when ``with_directions`` is false (lon / lat go to the device)."""
    if interp_obstimes is None:
        earth_xyz = _body_xyz("earth", skycoord.obstime, ephemeris).flatten()
    else:
        earth_xyz = interp_bodypos("earth", skycoord.obstime.mjd, interp_obstimes, ephemeris)

    if obspos_isstr:
        if obspos == "semb-l2":
            obs_xyz = semb_l2(earth_xyz)
        elif obspos == "earth":
            obs_xyz = earth_xyz
        elif skycoord.obstime.size == 1:
            try:
                obs_xyz = _body_xyz(obspos, skycoord.obstime, ephemeris).flatten()
            except KeyError as error:
                valid = [*coords.solar_system_ephemeris.bodies, "semb-l2"]
                raise ValueError(
                    f"Invalid observer string: '{obspos}'. Valid observers are: {valid}") from error
        else:
            obs_xyz = interp_bodypos(obspos, skycoord.obstime.mjd, interp_obstimes, ephemeris)
    else:
        try:
            obs_xyz = obspos.to_value(units.AU)
        except units.UnitConversionError as error:
            raise units.UnitConversionError("The observer position must be in length units.") from error

    if skycoord.obstime.size == 1:
        obs_xyz = obs_xyz[:, np.newaxis]
    if earth_xyz.ndim == 1:
        earth_xyz = earth_xyz[:, np.newaxis]

    return earth_xyz, obs_xyz, (sky_unit_vectors(skycoord) if with_directions else None)


def as_mjy_per_sr(emission):
    return emission << (units.MJy / units.sr)


def device_ephemeris(skycoord, obspos, interp_obstimes, ephemeris, device, with_directions=True):
    """Hourly Earth (and observer-body) knots as a device-resident spline
    (:class:`zodipy_b200.engine.DeviceEphemeris`) for time-ordered data with a string ``obspos``.

    Same knots as the host path (``arrange_obstimes`` + ``get_body``, ``bodies.py:16-35``); the
    per-sample interpolation and the SEMB-L2 scaling then happen in the kernel prologue instead of
    ``CubicSpline(...)(obstimes)`` on the host.  Returns (ephemeris, observer_mode, ecliptic unit
    vectors, sample times in MJD).
    """
    from .engine import DeviceEphemeris, MultiDeviceEphemeris

    knots_mjd = interp_obstimes.mjd
    earth_knots = _body_xyz("earth", interp_obstimes, ephemeris)
    obs_knots, mode = None, obspos
    if obspos not in ("earth", "semb-l2"):
        try:
            obs_knots = _body_xyz(obspos, interp_obstimes, ephemeris)
        except KeyError as error:
            valid = [*coords.solar_system_ephemeris.bodies, "semb-l2"]
            raise ValueError(f"Invalid observer string: '{obspos}'. Valid observers are: {valid}") from error
        mode = "knots"
    # np.arange(t0, t1 + dt, dt) fills t0 + k * delta with delta = (t0 + dt) - t0 (NOT exactly dt):
    # use the array's own spacing so the device knots equal the reference's knot times bit for bit
    delta = float(knots_mjd[1] - knots_mjd[0])
    devices = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
    if len(devices) > 1:
        eph = MultiDeviceEphemeris(float(knots_mjd[0]), delta, earth_knots, obs_knots, devices)
    else:
        eph = DeviceEphemeris(float(knots_mjd[0]), delta, earth_knots, obs_knots, device=devices[0])
    u_xyz = sky_unit_vectors(skycoord) if with_directions else None
    return eph, mode, u_xyz, np.ascontiguousarray(skycoord.obstime.mjd, dtype=np.float64)
