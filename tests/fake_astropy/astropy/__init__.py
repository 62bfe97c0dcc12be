"""Minimal duck-typed stand-in for the parts of Astropy that zodipy_b200's host layer touches.

TEST INFRASTRUCTURE ONLY: Astropy is not installed in the build / GPU image, so the tests put this
directory on sys.path to exercise ``Model.evaluate`` (``zodipy_b200/model.py``, ``astro.py``).  It is
NOT an ephemeris: Earth follows a low-precision analytic orbit (Meeus), other bodies circular ones.
"""
from . import coordinates, time, units  # noqa: F401

__version__ = "0.0-standin"
