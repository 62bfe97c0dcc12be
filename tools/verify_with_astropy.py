#!/usr/bin/env python
"""Checks that need REAL Astropy (absent from the build image and the GPU box): run once on a machine
that has astropy + a CUDA device + this repository's built library.

    python tools/verify_with_astropy.py [--reference]     # --reference: also import zodipy and compare

1. Documented known answers of the reference (README.md:36-48, docs/usage.md:163-203), reproduced through
   ``zodipy_b200.Model.evaluate(SkyCoord)``:
       25 um, galactic (40, 60) deg, TOD 2022-01-01 12:00-12:02  ->  [27.52410841 27.66572294 27.81251906]
       25 um, galactic (40, 60) deg, 2020-01-01                  ->  25.08189292
       same, obspos="mars"                                       ->  8.36985535
       same, obspos=[0.87, -0.53, 0.001] AU                      ->  20.37750965
2. ``sky_rotation="device"`` (angles + one 3x3 rotation, trigonometry in the kernel prologue) against
   ``sky_rotation="host"`` (Astropy transforms every coordinate, as the reference does) for ICRS, Galactic,
   FK5, BarycentricMeanEcliptic and HeliocentricMeanEcliptic coordinates, scalar and per-sample obstime,
   and that the device path is actually taken (``astro.sky_lonlat_rotation`` returns a rotation).
3. ``tod_ephemeris="device"`` against ``"host"`` for obspos in earth / semb-l2 / mars.
Exit status 0 = all checks passed.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FAILED = []


def check(label, got, want, rtol):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = float(np.max(np.abs(got - want) / np.abs(want)))
    ok = err <= rtol
    print(f"{'PASS' if ok else 'FAIL'}  {label}: max rel diff {err:.3e} (tolerance {rtol:g})")
    if not ok:
        FAILED.append(label)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", action="store_true", help="also run the reference (import zodipy) side by side")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    args = ap.parse_args()
    import astropy.coordinates as coords
    import astropy.units as u
    from astropy.time import Time, TimeDelta

    import zodipy_b200 as zp
    from zodipy_b200 import astro

    tol = 1e-8 if args.precision == "fp64" else 1e-5  # the documented values carry 8 decimals
    kw = {"precision": args.precision}

    # ---- 1. documented known answers ----
    model = zp.Model(25 * u.micron, **kw)
    tod = coords.SkyCoord([40, 41, 42] * u.deg, [60, 59, 58] * u.deg, frame="galactic",
                          obstime=Time("2022-01-01 12:00:00") + TimeDelta(np.arange(3) * 60.0, format="sec"))
    print("README TOD example:", np.asarray(model.evaluate(tod)))
    single = coords.SkyCoord(40 * u.deg, 60 * u.deg, frame="galactic", obstime=Time("2020-01-01"))
    check("docs/usage.md:174  25 um galactic (40, 60) 2020-01-01", model.evaluate(single), [25.08189292], max(tol, 5e-9))
    check("docs/usage.md:199  obspos='mars'", model.evaluate(single, obspos="mars"), [8.36985535], max(tol, 5e-9))
    check("docs/usage.md:203  obspos=[0.87, -0.53, 0.001] AU",
          model.evaluate(single, obspos=[0.87, -0.53, 0.001] * u.AU), [20.37750965], max(tol, 5e-9))

    # ---- 2. device rotation vs Astropy's transform_to ----
    rng = np.random.default_rng(0)
    n = 5000
    lon, lat = rng.uniform(0, 360, n) * u.deg, np.degrees(np.arcsin(rng.uniform(-1, 1, n))) * u.deg
    t_one = Time("2021-03-04T05:06:07")
    t_many = Time(59215.0 + np.sort(rng.uniform(0, 40, n)), format="mjd")
    frames = {"icrs": coords.ICRS, "galactic": coords.Galactic, "fk5": coords.FK5,
              "barycentricmeanecliptic": coords.BarycentricMeanEcliptic,
              "heliocentricmeanecliptic": coords.HeliocentricMeanEcliptic}
    for name, frame in frames.items():
        for label, obstime in (("scalar obstime", t_one), ("per-sample obstime", t_many)):
            try:
                sc = coords.SkyCoord(lon, lat, frame=frame, obstime=obstime)
            except Exception as error:  # noqa: BLE001 - report and continue with the other frames
                print(f"SKIP  {name}, {label}: SkyCoord construction failed ({error})")
                continue
            rot = astro.sky_lonlat_rotation(sc)
            print(f"      {name}, {label}: device rotation {'taken' if rot is not None else 'NOT taken (host fallback)'}")
            if rot is not None:
                lon_r, lat_r, r = rot
                vec = r @ np.array([np.cos(lat_r) * np.cos(lon_r), np.cos(lat_r) * np.sin(lon_r), np.sin(lat_r)])
                want = astro.sky_unit_vectors(sc)
                ok = np.allclose(vec, want, rtol=0, atol=1e-12)
                print(f"{'PASS' if ok else 'FAIL'}  {name}, {label}: R @ (lon, lat) == transform_to(...).cartesian "
                      f"(max abs diff {np.max(np.abs(vec - want)):.2e})")
                if not ok:
                    FAILED.append(f"rotation {name} {label}")
            dev = zp.Model(25 * u.micron, sky_rotation="device", **kw).evaluate(sc)
            host = zp.Model(25 * u.micron, sky_rotation="host", **kw).evaluate(sc)
            check(f"{name}, {label}: sky_rotation device vs host", dev, host, 1e-11 if args.precision == "fp64" else 3e-6)

    # ---- 3. device ephemeris spline vs host CubicSpline ----
    sc = coords.SkyCoord(lon, lat, frame="galactic", obstime=t_many)
    for obspos in ("earth", "semb-l2", "mars"):
        dev = zp.Model(25 * u.micron, tod_ephemeris="device", **kw).evaluate(sc, obspos=obspos)
        host = zp.Model(25 * u.micron, tod_ephemeris="host", **kw).evaluate(sc, obspos=obspos)
        check(f"tod_ephemeris device vs host, obspos={obspos}", dev, host, 1e-10 if args.precision == "fp64" else 3e-6)

    if args.reference:
        import zodipy

        ref = zodipy.Model(25 * u.micron)
        for label, coord, obspos in (("single", single, "earth"), ("tod", tod, "earth"), ("galactic map", sc, "semb-l2")):
            check(f"vs zodipy.Model ({label}, obspos={obspos})", model.evaluate(coord, obspos=obspos),
                  ref.evaluate(coord, obspos=obspos), 1e-10 if args.precision == "fp64" else 1e-5)

    print("ALL CHECKS PASSED" if not FAILED else f"{len(FAILED)} CHECK(S) FAILED: {FAILED}")
    return 1 if FAILED else 0


if __name__ == "__main__":
    sys.exit(main())
