import datetime as _dt

import numpy as np


def _iso_to_mjd(text):
    text = text.replace("T", " ")
    fmt = "%Y-%m-%d %H:%M:%S" if ":" in text else "%Y-%m-%d"
    t = _dt.datetime.strptime(text.split(".")[0], fmt)
    return (t - _dt.datetime(1858, 11, 17)).total_seconds() / 86400.0


class TimeDelta:
    def __init__(self, value, format="jd"):
        self.days = np.asarray(value, dtype=np.float64)


class Time:
    def __init__(self, value, format=None, scale="utc"):
        if isinstance(value, Time):
            self.mjd = value.mjd
        elif isinstance(value, str):
            self.mjd = np.float64(_iso_to_mjd(value))
        elif isinstance(value, (list, tuple)) and value and isinstance(value[0], str):
            self.mjd = np.array([_iso_to_mjd(v) for v in value])
        else:
            v = np.asarray(value, dtype=np.float64)
            self.mjd = v if format in (None, "mjd") else v - 2400000.5
        self.mjd = np.asarray(self.mjd, dtype=np.float64)

    @property
    def size(self):
        return int(self.mjd.size)

    @property
    def shape(self):
        return self.mjd.shape

    @property
    def jd(self):
        return self.mjd + 2400000.5

    def __getitem__(self, item):
        return Time(self.mjd[item], format="mjd")

    def __add__(self, delta):
        return Time(self.mjd + delta.days, format="mjd")

    def __len__(self):
        return len(self.mjd)
