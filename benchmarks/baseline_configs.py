#!/usr/bin/env python
"""Kernel-resident throughput + oracle parity for the five BASELINE.json configurations.

Informational companion of bench.py (which times only the headline configuration): for each
config, inputs are generated on the device, the fp32 and fp64 modes are timed with CUDA events and a
random subset is checked against the CPU oracle.  Prints one JSON object per config.

    python benchmarks/baseline_configs.py [--max-n 2e8] [--skip-fp64-above 6e7]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import zodi_oracle as oracle  # noqa: E402
import zodipy_b200 as zp  # noqa: E402
from zodipy_b200 import engine, healpix  # noqa: E402

EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])


def healpix_dirs(nside, dev):
    u = torch.empty((3, 12 * nside * nside), dtype=torch.float64, device=dev)
    _cabi = engine._cabi
    _cabi.check(_cabi.load().zodi_healpix_vectors(dev.index, nside, 0, 0, u.shape[1], None, u.data_ptr(),
                                                  u.shape[1], _cabi.MEM_DEVICE, None))
    torch.cuda.synchronize()
    return u


def tod_inputs(n, dev):
    """Synthetic year of time-ordered data: spin-scan-like pointing, Earth on a slightly eccentric
    orbit, observer = SEMB-L2-like scaling of Earth plus a small halo orbit (per-sample arrays)."""
    i = torch.arange(n, dtype=torch.float64, device=dev)
    t = i / n  # years
    lon_e = 2 * np.pi * t + 1.7
    r_e = 1.0 - 0.0167 * torch.cos(lon_e - 1.8)
    earth = torch.stack([r_e * torch.cos(lon_e), r_e * torch.sin(lon_e), 1e-5 * torch.sin(3 * lon_e)])
    obs = earth * 1.01 + torch.stack([torch.zeros_like(t), torch.zeros_like(t), 0.002 * torch.cos(40 * lon_e)])
    spin = 2 * np.pi * 525600.0 * t  # one rotation per minute
    colat = np.radians(85.0) + np.radians(7.5) * torch.sin(2 * np.pi * 8760.0 * t)
    # pointing = rotate (colat, spin) about the anti-sun axis
    ax = -earth / torch.linalg.vector_norm(earth, dim=0, keepdim=True)
    z = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64, device=dev).reshape(3, 1).expand(3, n)
    e1 = torch.linalg.cross(ax, z, dim=0)
    e1 = e1 / torch.linalg.vector_norm(e1, dim=0, keepdim=True)
    e2 = torch.linalg.cross(ax, e1, dim=0)
    u = torch.cos(colat) * ax + torch.sin(colat) * (torch.cos(spin) * e1 + torch.sin(spin) * e2)
    u = u / torch.linalg.vector_norm(u, dim=0, keepdim=True)
    return u.contiguous(), obs.contiguous(), earth.contiguous()


def time_mode(dm, u, obs, earth, precision, out_dtype, reps):
    out = dm.evaluate(u, obs, earth, precision=precision, out_dtype=out_dtype)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dm.evaluate(u, obs, earth, precision=precision, out=out, out_dtype=out_dtype)
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def run(label, model, u, obs, earth, skip_fp64_above):
    dm = model.device_model
    n = u.shape[1]
    units = n * model.ncomps * len(model.spec["points"])
    sel = np.sort(np.random.default_rng(2).choice(n, 1500, replace=False))
    sel_t = torch.as_tensor(sel, device=u.device)
    per_sample = obs.shape[1] == n
    u_s = u[:, sel_t].cpu().numpy()
    obs_s = obs[:, sel_t].cpu().numpy() if per_sample else obs.cpu().numpy()
    earth_s = earth[:, sel_t].cpu().numpy() if per_sample else earth.cpu().numpy()
    # the oracle derives its early-out flags from the subset's observers; they equal the global
    # ones here because every synthetic observer stays inside all outer cutoffs
    ref = oracle.evaluate(model.spec, u_s, obs_s, earth_s).sum(axis=0)
    res = {"config": label, "n_los": n, "ncomps": model.ncomps, "evaluations": units,
           "kernel": {p: dm.kernel_name_for(n, p) for p in ("fp32", "fp64")}}
    for precision, out_dtype, tol in (("fp32", np.float32, 1e-5), ("fp64", np.float64, 1e-10)):
        if precision == "fp64" and n > skip_fp64_above:
            continue
        reps = 5 if precision == "fp32" else 2
        out, ms = time_mode(dm, u, obs, earth, precision, out_dtype, reps)
        got = out[sel_t].double().cpu().numpy()
        err = float(np.max(np.abs(got - ref) / np.abs(ref)))
        res[precision] = {"ms": ms, "evals_per_s": units / (ms * 1e-3), "los_per_s": n / (ms * 1e-3),
                          "max_rel_err_vs_oracle": err, "tolerance": tol, "ok": bool(err <= tol)}
    print(json.dumps(res), flush=True)
    return res


def tod_e2e(dev, n=20_000_000):
    """Time-ordered data END TO END from pinned host memory: (a) the reference's array seam with
    per-sample observer/Earth arrays (72 B per sample up), (b) on-device ephemeris splines
    (pointing + time, 32 B per sample up), (c) the same with the pointing as longitude / latitude
    (24 B per sample up; unit vectors formed in the kernel prologue)."""
    import time

    from scipy.interpolate import CubicSpline

    t0, dt = 59215.0, 1.0 / 24.0
    n_knots = 366 * 24
    tk = t0 + dt * np.arange(n_knots)
    lon = 2 * np.pi * (tk - t0) / 365.25 + 1.7
    r = 1.0 - 0.0167 * np.cos(lon - 1.8)
    earth_knots = np.array([r * np.cos(lon), r * np.sin(lon), 1e-5 * np.sin(3 * lon)])
    t = np.linspace(t0, t0 + 365.0, n)
    u_dev, _, _ = tod_inputs(n, dev)
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    u = pin(u_dev.cpu().numpy())
    del u_dev
    tic = time.perf_counter()
    earth = CubicSpline(tk, earth_knots, axis=-1)(t)  # what the reference does per sample on the host
    host_interp_s = time.perf_counter() - tic
    earth, t_p = pin(earth), pin(t)
    lon_p, lat_p = pin(np.arctan2(u[1], u[0])), pin(np.arcsin(np.clip(u[2], -1.0, 1.0)))
    out = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
    model = zp.Model(zp.Quantity(25.0, "um"), precision="fp32")
    eph = engine.DeviceEphemeris(t0, dt, earth_knots)
    res = {"config": f"4e: TOD e2e {n:.0e} samples from pinned host memory, dirbe 25um, observer=earth",
           "host_cubicspline_seconds_not_timed_below": host_interp_s}
    variants = (
        ("array_seam_72B_per_sample", lambda: model.evaluate_xyz(u, earth, earth, out=out, out_dtype=np.float32)),
        ("device_ephemeris_32B_per_sample", lambda: model.evaluate_tod_xyz(u, t_p, eph, out=out, out_dtype=np.float32)),
        ("device_ephemeris_lonlat_24B_per_sample",
         lambda: model.evaluate_lonlat(lon_p, lat_p, ephemeris=eph, obstime=t_p, out=out, out_dtype=np.float32)),
        ("device_ephemeris_times_sent_twice_40B", lambda: model.evaluate_tod_xyz(u, t_p, eph, out=out, out_dtype=np.float32)))
    best = {label: float("inf") for label, _ in variants}
    sums = {}
    for rep in range(6):  # interleaved, best of 5 after one warm-up round (wall clock: allocator / OS noise)
        for label, call in variants:
            os.environ.pop("ZODI_TOD_EXPLICIT_OBSTIME", None)
            if label.endswith("40B"):
                os.environ["ZODI_TOD_EXPLICIT_OBSTIME"] = "1"
            torch.cuda.synchronize()
            tic = time.perf_counter()
            call()
            torch.cuda.synchronize()
            if rep:
                best[label] = min(best[label], (time.perf_counter() - tic) * 1e3)
            sums[label] = float(out.sum(dtype=np.float64))
    os.environ.pop("ZODI_TOD_EXPLICIT_OBSTIME", None)
    for label, _ in variants:
        res[label] = {"ms": best[label], "evals_per_s": n * 6 * 50 / (best[label] * 1e-3), "result_sum": sums[label]}
    print(json.dumps(res), flush=True)


def multiband(dev):
    """All bands of a model for the same pointings: MultiBandModel vs the per-band loop."""
    for name, xs, unit, nside in (("dirbe", [1.25, 2.2, 3.5, 4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0], "um", 512),
                                  ("dirbe", [4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0], "um", 512),
                                  ("planck18", [100.0, 143.0, 217.0, 353.0, 545.0, 857.0], "GHz", 1024)):
        u = healpix_dirs(nside, dev)
        earth = torch.as_tensor(EARTH, device=dev)
        mb = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name, precision="fp32")
        res = {"config": f"multiband: {name} {len(xs)} bands nside={nside}", "n_los": u.shape[1], "n_bands": len(xs)}
        out = mb.evaluate_xyz(u, earth, out_dtype=np.float32)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            mb.evaluate_xyz(u, earth, out=out, out_dtype=np.float32)
        b.record()
        torch.cuda.synchronize()
        ms_mb = a.elapsed_time(b) / 5
        # the scalar multi-band kernel (one line of sight per thread) on the same input
        os.environ["ZODI_NO_X2"] = "1"
        mbs = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name, precision="fp32")
        out_s = mbs.evaluate_xyz(u, earth, out_dtype=np.float32)
        os.environ.pop("ZODI_NO_X2")
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            mbs.evaluate_xyz(u, earth, out=out_s, out_dtype=np.float32)
        b.record()
        torch.cuda.synchronize()
        res["scalar_multiband_ms"] = a.elapsed_time(b) / 5
        res["max_rel_diff_vs_scalar_multiband"] = float(((out - out_s).abs() / out_s.abs()).max())
        singles = [torch.empty(u.shape[1], dtype=torch.float32, device=dev) for _ in xs]
        for m, o in zip(mb.bands, singles):
            m.evaluate_xyz(u, earth, out=o, out_dtype=np.float32)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            for m, o in zip(mb.bands, singles):
                m.evaluate_xyz(u, earth, out=o, out_dtype=np.float32)
        b.record()
        torch.cuda.synchronize()
        ms_loop = a.elapsed_time(b) / 5
        diff = max(float(((out[i] - singles[i]).abs() / singles[i].abs()).max()) for i in range(len(xs)))
        band_evals = u.shape[1] * mb.bands[0].ncomps * 50 * len(xs)
        res.update({"kernel": mb.device_model.kernel_name_for(u.shape[1], "fp32"),
                    "multiband_ms": ms_mb, "per_band_loop_ms": ms_loop, "speedup": ms_loop / ms_mb,
                    "band_evals_per_s": band_evals / (ms_mb * 1e-3), "max_rel_diff_vs_single_band": diff})
        print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=float, default=2.1e8)
    ap.add_argument("--skip-fp64-above", type=float, default=6e7)
    ap.add_argument("--scatter-only", action="store_true", help="only the 1.25 um scattering case, nside 1024")
    ap.add_argument("--tod-only", action="store_true", help="only the end-to-end time-ordered case")
    ap.add_argument("--multiband-only", action="store_true", help="only the multi-band cases")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    earth = torch.as_tensor(EARTH, device=dev)
    Q = zp.Quantity
    if args.tod_only:
        tod_e2e(dev)
        return
    if args.multiband_only:
        multiband(dev)
        return
    if args.scatter_only:
        run("extra: dirbe 1.25um (scattering) nside=1024", zp.Model(Q(1.25, "um")), healpix_dirs(1024, dev), earth,
            earth, args.skip_fp64_above)
        return

    run("1: dirbe 25um nside=64", zp.Model(Q(25.0, "um")), healpix_dirs(64, dev), earth, earth, args.skip_fp64_above)
    x = np.linspace(9.0, 15.0, 10)
    w = np.exp(-0.5 * ((x - 12.0) / 1.5) ** 2)
    run("2: dirbe 12um 10-sample bandpass nside=512", zp.Model(Q(x, "um"), weights=w), healpix_dirs(512, dev),
        earth, earth, args.skip_fp64_above)
    run("3: planck18 857GHz nside=2048", zp.Model(Q(857.0, "GHz"), name="planck18"), healpix_dirs(2048, dev),
        earth, earth, args.skip_fp64_above)
    n_tod = int(min(1e8, args.max_n))
    u, obs, ear = tod_inputs(n_tod, dev)
    run(f"4: TOD {n_tod:.0e} samples dirbe 25um per-sample obs/earth", zp.Model(Q(25.0, "um")), u, obs, ear,
        args.skip_fp64_above)
    del u, obs, ear
    torch.cuda.empty_cache()
    if args.max_n >= 12 * 4096 * 4096:
        run("5: planck13 545GHz nside=4096", zp.Model(Q(545.0, "GHz"), name="planck13"), healpix_dirs(4096, dev),
            earth, earth, 3e8)
    tod_e2e(dev)
    multiband(dev)
    # extra: scattering branch and the generic kernel (RRM)
    run("extra: dirbe 1.25um (scattering) nside=512", zp.Model(Q(1.25, "um")), healpix_dirs(512, dev), earth, earth,
        args.skip_fp64_above)
    run("extra: rrm-experimental 25um nside=256 (generic kernel)", zp.Model(Q(25.0, "um"), name="rrm-experimental"),
        healpix_dirs(256, dev), earth, earth, args.skip_fp64_above)


if __name__ == "__main__":
    main()
