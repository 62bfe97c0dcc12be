import numpy as np

from . import units as u


class _Frame:
    name = ""


class BarycentricMeanEcliptic(_Frame):
    name = "barycentricmeanecliptic"


class HeliocentricMeanEcliptic(_Frame):
    name = "heliocentricmeanecliptic"


class ICRS(_Frame):
    name = "icrs"


class Galactic(_Frame):
    name = "galactic"


_EPS = np.radians(23.4392911)
_ICRS_TO_ECL = np.array([[1, 0, 0], [0, np.cos(_EPS), np.sin(_EPS)], [0, -np.sin(_EPS), np.cos(_EPS)]])
# ICRS -> Galactic (Hipparcos); stand-in precision is irrelevant, it only has to be a rotation
_ICRS_TO_GAL = np.array([[-0.0548755604, -0.8734370902, -0.4838350155],
                         [0.4941094279, -0.4448296300, 0.7469822445],
                         [-0.8676661490, -0.1980763734, 0.4559837762]])
_U, _, _VT = np.linalg.svd(_ICRS_TO_GAL)
_ICRS_TO_GAL = _U @ _VT  # nearest exact rotation to the 10-digit table
_TO_ECL = {"barycentricmeanecliptic": np.eye(3), "heliocentricmeanecliptic": np.eye(3),
           "icrs": _ICRS_TO_ECL, "galactic": _ICRS_TO_ECL @ _ICRS_TO_GAL.T}


def _frame_name(frame):
    if frame is None:
        return "icrs"
    if isinstance(frame, str):
        return frame.lower()
    return frame.name


class _Cartesian:
    def __init__(self, xyz):
        self.xyz = u.Quantity(xyz, u.AU)


class UnitSphericalRepresentation:
    def __init__(self, lon_rad, lat_rad):
        self.lon, self.lat = u.Quantity(lon_rad, u.rad), u.Quantity(lat_rad, u.rad)


class _FrameInstance(_Frame):
    def __init__(self, name):
        self.name = name

    def replicate_without_data(self):
        return _FrameInstance(self.name)


class SkyCoord:
    def __init__(self, lon, lat=None, unit=None, frame=None, obstime=None, _xyz=None, _scalar=None):
        self.frame_name = _frame_name(frame)
        self.frame = _FrameInstance(self.frame_name)
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
        rot = _TO_ECL[target].T @ _TO_ECL[self.frame_name]
        xyz = rot @ self._xyz.reshape(3, -1)
        return SkyCoord(None, frame=target, obstime=self.obstime, _xyz=xyz.reshape(self._xyz.shape),
                        _scalar=self.isscalar)

    def __getitem__(self, item):
        out = SkyCoord(None, frame=self.frame_name, obstime=self.obstime, _xyz=self._xyz[:, item],
                       _scalar=False)
        if isinstance(self.data, UnitSphericalRepresentation):
            out.data = UnitSphericalRepresentation(np.asarray(self.data.lon)[item], np.asarray(self.data.lat)[item])
        return out


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
