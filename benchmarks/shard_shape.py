#!/usr/bin/env python
"""Launch shape of the packed kernel on ONE rank's share of a map split over several GPUs.

At 8 GPUs a rank integrates 1/8 of the nside-2048 map (6.3 M lines of sight, 16.6 waves of 128-thread CTAs):
the last, partly filled wave is ~3 % of the launch.  More lanes per pair of lines of sight halve the CTA's
run time (twice as many CTAs) at some cost in per-line prologue work.  This script times the block-cyclic
share of rank 0 of `--ranks` ranks on one GPU (no peer stores) for every lane count.  One JSON line each.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "benchmarks")]
import zodipy_b200 as zp  # noqa: E402
from n_sweep import EARTH, directions, time_launches  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nside", type=int, default=2048)
    ap.add_argument("--ranks", default="8,4,2")
    ap.add_argument("--block", type=int, default=16384)
    ap.add_argument("--model", default="planck18")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    x, unit = {"planck18": (857.0, "GHz"), "dirbe": (25.0, "um")}[args.model]
    model = zp.Model(zp.Quantity(x, unit), name=args.model, precision="fp32")
    dm = model.device_model
    flags = dm.outside_flags(EARTH)
    obs = torch.as_tensor(EARTH, device=dev)
    u_all = directions(args.nside, dev)
    n_all = u_all.shape[1]
    for ranks in (int(r) for r in args.ranks.split(",")):
        idx = torch.arange(n_all, device=dev).view(-1, args.block)[0::ranks].reshape(-1)  # rank 0's blocks
        u = u_all[:, idx].contiguous()
        n = u.shape[1]
        out = torch.empty(n, dtype=torch.float32, device=dev)
        units = n * model.ncomps * len(model.spec["points"])
        for lanes in (0, 1, 2, 4):
            if lanes:
                os.environ["ZODI_X2_LANES"] = str(lanes)
            else:
                os.environ.pop("ZODI_X2_LANES", None)

            def call():
                dm.evaluate(u, obs, obs, precision="fp32", out=out, out_dtype=np.float32, outside_flags=flags)

            ms, how = time_launches(call, per_graph=8, replays=8)
            print(json.dumps({"model": args.model, "nside": args.nside, "ranks": ranks, "n_los": n,
                              "lanes": lanes or "auto", "ms": ms, "ideal_ms_from_full_map": None,
                              "evals_per_s": units / (ms * 1e-3), "timing": how,
                              "checksum": float(out.double().sum().item())}), flush=True)
        os.environ.pop("ZODI_X2_LANES", None)
        del u, out


if __name__ == "__main__":
    main()
