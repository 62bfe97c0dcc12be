"""End-to-end (pinned host arrays in, host map out) rate of ONE process driving k GPUs through
Model(devices=[0..k-1]) - the GPU counterpart of the reference's `nprocesses` pool - for the
bench.py workload (planck18 857 GHz, nside-2048 map, fp32), plus the box's host -> device copy ceiling
for k concurrent links (plain cudaMemcpyAsync, no kernels): slices of ONE pinned buffer vs one pinned
buffer per GPU.  One JSON line per k and entry."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zodipy_b200 as zp  # noqa: E402
from zodipy_b200 import healpix  # noqa: E402

EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])


def h2d_ceiling(k, nbytes_per_gpu=1 << 28, reps=4):
    """Aggregate H2D GB/s of k GPUs copying concurrently (one stream each, issued from this thread)."""
    n = nbytes_per_gpu // 8
    res = {}
    shared = torch.empty(k * n, dtype=torch.float64).pin_memory()
    separate = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(k)]
    dst = [torch.empty(n, dtype=torch.float64, device=f"cuda:{d}") for d in range(k)]
    for label, srcs in (("one_pinned_buffer", [shared[d * n:(d + 1) * n] for d in range(k)]),
                        ("pinned_buffer_per_gpu", separate)):
        best = float("inf")
        for _ in range(reps + 1):
            for d in range(k):
                torch.cuda.synchronize(d)
            tic = time.perf_counter()
            for d in range(k):
                with torch.cuda.device(d):
                    dst[d].copy_(srcs[d], non_blocking=True)
            for d in range(k):
                torch.cuda.synchronize(d)
            best = min(best, time.perf_counter() - tic)
        res[label] = k * nbytes_per_gpu / best / 1e9
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nside", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    npix = healpix.nside2npix(args.nside)
    u = torch.empty((3, npix), dtype=torch.float64).pin_memory()
    healpix.full_sky_vectors(args.nside, 0, npix, out=u.numpy())
    u = u.numpy()
    lon = torch.empty(npix, dtype=torch.float64).pin_memory().numpy()
    lat = torch.empty(npix, dtype=torch.float64).pin_memory().numpy()
    np.arctan2(u[1], u[0], out=lon)
    np.arcsin(np.clip(u[2], -1.0, 1.0), out=lat)
    out = torch.empty(npix, dtype=torch.float32).pin_memory().numpy()
    ref = None
    ngpu = torch.cuda.device_count()
    for k in sorted({1, 2, 4, 8, ngpu}):
        if k > ngpu:
            continue
        model = zp.Model(zp.Quantity(857.0, "GHz"), name="planck18", precision="fp32", devices=list(range(k)))
        print(json.dumps({"gpus_one_process": k, "entry": "h2d_ceiling_GBps", **h2d_ceiling(k)}), flush=True)
        entries = (("evaluate_xyz", lambda: model.evaluate_xyz(u, EARTH, EARTH, out=out, out_dtype=np.float32), 24),
                   ("evaluate_lonlat", lambda: model.evaluate_lonlat(lon, lat, EARTH, EARTH, out=out,
                                                                     out_dtype=np.float32), 16),
                   ("evaluate_healpix", lambda: model.evaluate_healpix(args.nside, EARTH, out=out,
                                                                       out_dtype=np.float32), 0))
        for name, call, bytes_in in entries:
            for _ in range(2):
                call()
            best = float("inf")
            for _ in range(args.reps):
                for d in range(k):
                    torch.cuda.synchronize(d)
                tic = time.perf_counter()
                call()
                best = min(best, time.perf_counter() - tic)
            if ref is None:
                ref = out.copy()
            print(json.dumps({"gpus_one_process": k, "entry": name, "nside": args.nside, "ms_best": best * 1e3,
                              "evals_per_s": npix * 4 * 50 / best, "h2d_GBps": npix * bytes_in / best / 1e9,
                              "d2h_GBps": npix * 4 / best / 1e9,
                              "identical_to_one_gpu_xyz": bool(np.array_equal(out, ref))}), flush=True)


if __name__ == "__main__":
    main()
