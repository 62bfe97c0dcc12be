"""A/B timing of kernel variants: device-resident HEALPix maps (directions generated in the kernel
prologue), CUDA events, best and median of `reps` launches.  Select the library build with
ZODI_B200_LIB=<path to another libzodi_b200.so>.  One JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zodipy_b200 as zp  # noqa: E402

EARTH = np.array([-0.3919640703, 0.9020953332, 0.0])
CASES = (("planck18", 857.0, "GHz", 2048, "fp32"), ("dirbe", 25.0, "um", 1024, "fp32"),
         ("planck13", 545.0, "GHz", 1024, "fp32"), ("planck18", 857.0, "GHz", 1024, "fp64"))


def main():
    reps = int(os.environ.get("AB_REPS", 15))
    only = os.environ.get("AB_ONLY", "")  # substring filter on the case label, e.g. AB_ONLY=fp64
    for name, x, unit, nside, precision in CASES:
        if only and only not in f"{name} {x}{unit} nside{nside} {precision}":
            continue
        model = zp.Model(zp.Quantity(x, unit), name=name, precision=precision)
        out = None
        ts = []
        for i in range(reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = model.evaluate_healpix(nside, EARTH, out=out, out_dtype=np.float32, device_out=True)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        n = 12 * nside * nside * model.ncomps * 50
        print(json.dumps({"lib": os.environ.get("ZODI_B200_LIB", "default"), "case": f"{name} {x}{unit} nside{nside} {precision}",
                          "ms_min": min(ts), "ms_median": float(np.median(ts)), "evals_per_s": n / (min(ts) * 1e-3),
                          "checksum": float(out.double().sum().item())}), flush=True)


if __name__ == "__main__":
    main()
