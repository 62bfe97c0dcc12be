import numpy as np

from . import units as u


class ConvertError(Exception):
    pass


class _Frame:
    """Data-less frame.  ``frame_attributes`` mirrors Astropy's mapping of attribute names; the stand-in
    frames that list ``obstime`` there receive the SkyCoord's obstime (like Astropy's
    HeliocentricMeanEcliptic / GCRS do), possibly one value per coordinate."""
    name = ""
    frame_attributes = {}

    def __init__(self, **attrs):
        unknown = set(attrs) - set(self.frame_attributes)
        if unknown:
            raise TypeError(f"unexpected frame attributes {sorted(unknown)}")
        for key in self.frame_attributes:
            setattr(self, key, attrs.get(key))

    def replicate_without_data(self):
        return type(self)(**{k: getattr(self, k) for k in self.frame_attributes})


class BarycentricMeanEcliptic(_Frame):
    name = "barycentricmeanecliptic"


class HeliocentricMeanEcliptic(_Frame):
    name = "heliocentricmeanecliptic"


class ICRS(_Frame):
    name = "icrs"


class Galactic(_Frame):
    name = "galactic"


class ObstimeEcliptic(_Frame):
    """Mean-ecliptic frame that carries ``obstime`` as a FRAME attribute (stand-in for Astropy frames such
    as HeliocentricMeanEcliptic whose instances hold the per-sample obstime of time-ordered data)."""
    name = "obstimeecliptic"
    frame_attributes = {"obstime": None}


class TimeDependentFrame(_Frame):
    """A frame whose orientation depends on its obstime attribute: NOT a fixed rotation."""
    name = "timedependent"
    frame_attributes = {"obstime": None}


_EPS = np.radians(23.4392911)
_ICRS_TO_ECL = np.array([[1, 0, 0], [0, np.cos(_EPS), np.sin(_EPS)], [0, -np.sin(_EPS), np.cos(_EPS)]])
# ICRS -> Galactic (Hipparcos); stand-in precision is irrelevant, it only has to be a rotation
_ICRS_TO_GAL = np.array([[-0.0548755604, -0.8734370902, -0.4838350155],
                         [0.4941094279, -0.4448296300, 0.7469822445],
                         [-0.8676661490, -0.1980763734, 0.4559837762]])
_U, _, _VT = np.linalg.svd(_ICRS_TO_GAL)
_ICRS_TO_GAL = _U @ _VT  # nearest exact rotation to the 10-digit table
_TO_ECL = {"barycentricmeanecliptic": np.eye(3), "heliocentricmeanecliptic": np.eye(3), "obstimeecliptic": np.eye(3),
           "icrs": _ICRS_TO_ECL, "galactic": _ICRS_TO_ECL @ _ICRS_TO_GAL.T}
_FRAME_CLASSES = {}


def _frame_name(frame):
    if frame is None:
        return "icrs"
    if isinstance(frame, str):
        return frame.lower()
    return frame.name


def _frame_instance(frame, obstime):
    """Frame object of a SkyCoord: an instance keeps its attributes, a class / name gets the SkyCoord's
    obstime if it declares that attribute."""
    if isinstance(frame, _Frame):
        return frame
    cls = frame if isinstance(frame, type) else _FRAME_CLASSES.get(_frame_name(frame), None)
    if cls is None:
        inst = _Frame()
        inst.name = _frame_name(frame)
        return inst
    return cls(**({"obstime": obstime} if "obstime" in cls.frame_attributes else {}))


class _Cartesian:
    def __init__(self, xyz):
        self.xyz = u.Quantity(xyz, u.AU)


class UnitSphericalRepresentation:
    def __init__(self, lon_rad, lat_rad):
        self.lon, self.lat = u.Quantity(lon_rad, u.rad), u.Quantity(lat_rad, u.rad)


def _rotation_to_ecliptic(frame, n):
    """(3, 3) or per-coordinate (n, 3, 3) rotation of the frame to the mean ecliptic."""
    if frame.name != "timedependent":
        return _TO_ECL[frame.name]
    mjd = np.atleast_1d(np.asarray(frame.obstime.mjd, dtype=np.float64))
    ang = 1e-3 * (mjd - 59000.0)  # slowly turning about z
    c, s_ = np.cos(ang), np.sin(ang)
    rot = np.zeros((mjd.size, 3, 3))
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1], rot[:, 2, 2] = c, -s_, s_, c, 1.0
    if mjd.size not in (1, n):
        raise ValueError("frame attribute obstime does not broadcast against the coordinates")
    return rot if mjd.size > 1 else rot[0]


class SkyCoord:
    def __init__(self, lon, lat=None, unit=None, frame=None, obstime=None, _xyz=None, _scalar=None):
        self.frame_name = _frame_name(frame)
        self.frame = _frame_instance(frame, obstime)
        if obstime is None and getattr(self.frame, "obstime", None) is not None:
            obstime = self.frame.obstime
        self.obstime = obstime
        if _xyz is not None:
            self._xyz, self.isscalar = _xyz, _scalar
            self.data = _Cartesian(_xyz)
            return
        if isinstance(lon, u.Quantity):  # angles carry their own unit
            lon_r, lat_r = np.asarray(lon.to_value(u.rad)), np.asarray(lat.to_value(u.rad))
        else:
            lon_r, lat_r = np.radians(np.asarray(lon, dtype=np.float64)), np.radians(np.asarray(lat, dtype=np.float64))
        self.isscalar = lon_r.ndim == 0
        self.data = UnitSphericalRepresentation(lon_r, lat_r)
        self._xyz = np.array([np.cos(lat_r) * np.cos(lon_r), np.cos(lat_r) * np.sin(lon_r), np.sin(lat_r)])

    @property
    def size(self):
        return 1 if self.isscalar else int(self._xyz.shape[1])

    @property
    def cartesian(self):
        return _Cartesian(self._xyz)

    def transform_to(self, frame):
        target = _frame_name(frame)
        flat = self._xyz.reshape(3, -1)
        n = flat.shape[1]
        fobs = getattr(self.frame, "obstime", None)
        if fobs is not None and np.size(fobs.mjd) not in (1, n):
            raise ValueError("frame attribute obstime does not broadcast against the coordinates")
        to_ecl = _rotation_to_ecliptic(self.frame, n)
        xyz = np.einsum("nij,jn->in", to_ecl, flat) if to_ecl.ndim == 3 else to_ecl @ flat
        xyz = _TO_ECL[target].T @ xyz
        return SkyCoord(None, frame=target, obstime=self.obstime, _xyz=xyz.reshape(self._xyz.shape),
                        _scalar=self.isscalar)

    def __getitem__(self, item):
        frame = self.frame
        fobs = getattr(frame, "obstime", None)
        if fobs is not None and np.size(fobs.mjd) > 1:
            frame = type(frame)(obstime=fobs[item])
        out = SkyCoord(None, frame=frame, obstime=self.obstime, _xyz=self._xyz[:, item],
                       _scalar=False)
        if isinstance(self.data, UnitSphericalRepresentation):
            out.data = UnitSphericalRepresentation(np.asarray(self.data.lon)[item], np.asarray(self.data.lat)[item])
        return out


_FRAME_CLASSES.update({c.name: c for c in (BarycentricMeanEcliptic, HeliocentricMeanEcliptic, ICRS, Galactic,
                                            ObstimeEcliptic, TimeDependentFrame)})


def _earth_xyz(mjd):
    """Low-precision heliocentric mean-ecliptic (J2000) Earth position (Meeus ch. 25) [AU]."""
    T = (np.asarray(mjd, dtype=np.float64) + 2400000.5 - 2451545.0) / 36525.0
    L0 = 280.46646 + 36000.76983 * T + 0.0003032 * T * T
    M = np.radians(357.52911 + 35999.05029 * T - 0.0001537 * T * T)
    e = 0.016708634 - 0.000042037 * T
    Cc = ((1.914602 - 0.004817 * T) * np.sin(M) + (0.019993 - 0.000101 * T) * np.sin(2 * M) + 0.000289 * np.sin(3 * M))
    nu = M + np.radians(Cc)
    R = 1.000001018 * (1 - e * e) / (1 + e * np.cos(nu))
    lon = np.radians(L0 + Cc + 180.0 - 1.396971 * T)
    return np.array([R * np.cos(lon), R * np.sin(lon), np.zeros_like(lon)])


_CIRCULAR = {"mars": (1.5237, 686.98, 0.9), "venus": (0.7233, 224.70, 2.1), "jupiter": (5.2029, 4332.6, 0.3),
             "moon": None}


class _SolarSystemEphemeris:
    bodies = ("earth", "sun", "moon", "mercury", "venus", "mars", "jupiter")


solar_system_ephemeris = _SolarSystemEphemeris()


def get_body(body, time, ephemeris=None):
    mjd = time.mjd
    if body == "earth":
        xyz = _earth_xyz(mjd)
    elif body == "moon":
        ang = 2 * np.pi * np.asarray(mjd) / 27.32
        xyz = _earth_xyz(mjd) + 0.00257 * np.array([np.cos(ang), np.sin(ang), 0.09 * np.sin(ang)])
    elif body in _CIRCULAR:
        a, period, phase = _CIRCULAR[body]
        ang = 2 * np.pi * np.asarray(mjd) / period + phase
        xyz = a * np.array([np.cos(ang), np.sin(ang), 0.03 * np.sin(ang)])
    else:
        raise KeyError(body)
    return SkyCoord(None, frame="heliocentricmeanecliptic", obstime=time, _xyz=xyz, _scalar=np.ndim(mjd) == 0)
