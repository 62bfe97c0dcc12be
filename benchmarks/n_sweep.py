#!/usr/bin/env python
"""Size sweep of the fp32 integrator kernels: HEALPix maps nside 32 ... 4096, device-resident inputs.

For every (model, nside) the kernel is timed for the launch shape the library picks and, with
``--shapes``, for every forced shape of the packed kernel (lanes per pair of lines of sight x CTA size;
env ZODI_X2_LANES / ZODI_X2_THREADS).  Small launches are shorter than the host's call overhead, so the
launches are captured into a CUDA graph (20 per replay) and the replay is timed with CUDA events: the
figure is device time per launch including the launch gap.  ``api_ms`` is the per-call time of the same
evaluation issued call by call through ``DeviceModel.evaluate`` (host overhead included).
One JSON line per measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import zodipy_b200 as zp  # noqa: E402
from zodipy_b200 import engine  # noqa: E402

EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])
MODELS = {"planck18": (857.0, "GHz"), "dirbe": (25.0, "um"), "planck13": (545.0, "GHz"),
          "rrm-experimental": (25.0, "um")}


def directions(nside, dev):
    n = 12 * nside * nside
    u = torch.empty((3, n), dtype=torch.float64, device=dev)
    cabi = engine._cabi
    cabi.check(cabi.load().zodi_healpix_vectors(dev.index, nside, 0, 0, n, None, u.data_ptr(), n,
                                                cabi.MEM_DEVICE, None))
    torch.cuda.synchronize()
    return u


def time_launches(call, per_graph=20, replays=5):
    """Device ms per launch: CUDA-graph replay of `per_graph` launches (falls back to a plain loop)."""
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    try:
        stream = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(per_graph):
                    call()
        run, how = graph.replay, "cuda graph"
    except Exception as exc:  # noqa: BLE001 - measurement helper: report and fall back
        print(f"graph capture failed ({exc}); timing a plain loop", file=sys.stderr)

        def run():
            for _ in range(per_graph):
                call()
        how = "loop"
    run()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(replays):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / per_graph)
    return best, how


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="planck18,dirbe")
    ap.add_argument("--nsides", default="32,64,128,256,512,1024,2048")
    ap.add_argument("--shapes", action="store_true", help="also force every packed-kernel shape")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--label", default=os.environ.get("ZODI_B200_LIB", "default"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    obs = torch.as_tensor(EARTH, device=dev)
    odt = np.float32 if args.precision == "fp32" else np.float64
    for name in args.models.split(","):
        x, unit = MODELS[name]
        model = zp.Model(zp.Quantity(x, unit), name=name, precision=args.precision)
        dm = model.device_model
        flags = dm.outside_flags(EARTH)
        for nside in (int(s) for s in args.nsides.split(",")):
            u = directions(nside, dev)
            n = u.shape[1]
            out = torch.empty(n, dtype=torch.float32 if args.precision == "fp32" else torch.float64, device=dev)
            units = n * model.ncomps * len(model.spec["points"])
            shapes = [(0, 0)]
            if args.shapes and "x2" in dm.kernel_name_for(n, args.precision):
                shapes += [(lanes, th) for th in (256, 128) for lanes in (1, 2, 4, 8)
                           if n * lanes <= 2 * 148 * 1280 * 16]
            for lanes, threads in shapes:
                for k, v in (("ZODI_X2_LANES", lanes), ("ZODI_X2_THREADS", threads)):
                    if v:
                        os.environ[k] = str(v)
                    else:
                        os.environ.pop(k, None)

                def call():
                    dm.evaluate(u, obs, obs, precision=args.precision, out=out, out_dtype=odt, outside_flags=flags)

                ms, how = time_launches(call, per_graph=20 if n < 4_000_000 else 4)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                reps = 50 if n < 4_000_000 else 5
                for _ in range(reps):
                    call()
                torch.cuda.synchronize()
                api_ms = (time.perf_counter() - t0) * 1e3 / reps
                print(json.dumps({"lib": args.label, "model": name, "precision": args.precision, "nside": nside,
                                  "n_los": n, "kernel": dm.kernel_name_for(n, args.precision),
                                  "lanes": lanes or "auto", "threads": threads or "auto", "ms": ms,
                                  "evals_per_s": units / (ms * 1e-3), "api_ms": api_ms, "timing": how,
                                  "checksum": float(out.double().sum().item())}), flush=True)
            os.environ.pop("ZODI_X2_LANES", None)
            os.environ.pop("ZODI_X2_THREADS", None)
            del u, out


if __name__ == "__main__":
    main()
