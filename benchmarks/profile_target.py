"""One model, a few full-sky HEALPix launches - the target process for `ncu` captures.

    ncu --set full --clock-control none --import-source on -k regex:zodi_los -s 2 -c 1 \
        -o gpurun_out/<name> python benchmarks/profile_target.py --x 1.25 --unit um --nside 1024

Pixel directions are generated in the kernel prologue (zodi_evaluate_healpix), the map stays on
the device; nothing here is timed.
"""
import argparse
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import zodipy_b200 as zp  # noqa: E402

EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="dirbe")
    ap.add_argument("--x", type=float, default=25.0)
    ap.add_argument("--unit", default="um")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--nside", type=int, default=1024)
    ap.add_argument("--launches", type=int, default=3)
    args = ap.parse_args()
    model = zp.Model(zp.Quantity(args.x, args.unit), name=args.name, precision=args.precision)
    dtype = np.float32 if args.precision == "fp32" else np.float64
    for _ in range(args.launches):
        model.evaluate_healpix(args.nside, EARTH, device_out=True, out_dtype=dtype)
    torch.cuda.synchronize()
    print(model.device_model.kernel_name_for(12 * args.nside**2, args.precision))


if __name__ == "__main__":
    main()
