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
    ap.add_argument("--arrays", action="store_true",
                    help="device-resident (3, N) unit vectors as input (bench.py's `value` path) instead of "
                         "directions generated in the kernel prologue")
    ap.add_argument("--bands", default="", help="comma-separated wavelengths / frequencies: MultiBandModel of "
                                                "these bands instead of the single band --x")
    args = ap.parse_args()
    if args.bands:
        mb = zp.MultiBandModel([zp.Quantity(float(b), args.unit) for b in args.bands.split(",")], name=args.name,
                               precision=args.precision)
        for _ in range(args.launches):
            mb.evaluate_healpix(args.nside, EARTH, device_out=True,
                                out_dtype=np.float32 if args.precision == "fp32" else np.float64)
        torch.cuda.synchronize()
        print(mb.device_model.kernel_name_for(12 * args.nside**2, args.precision))
        return
    model = zp.Model(zp.Quantity(args.x, args.unit), name=args.name, precision=args.precision)
    dtype = np.float32 if args.precision == "fp32" else np.float64
    if args.arrays:
        from zodipy_b200 import engine

        n = 12 * args.nside**2
        dev = torch.device("cuda", 0)
        u = torch.empty((3, n), dtype=torch.float64, device=dev)
        cabi = engine._cabi
        cabi.check(cabi.load().zodi_healpix_vectors(0, args.nside, 0, 0, n, None, u.data_ptr(), n, cabi.MEM_DEVICE, None))
        obs = torch.as_tensor(EARTH, device=dev)
        flags = model.device_model.outside_flags(EARTH)
        out = torch.empty(n, dtype=torch.float32 if args.precision == "fp32" else torch.float64, device=dev)
    for _ in range(args.launches):
        if args.arrays:
            model.device_model.evaluate(u, obs, obs, precision=args.precision, out=out, out_dtype=dtype,
                                        outside_flags=flags)
        else:
            model.evaluate_healpix(args.nside, EARTH, device_out=True, out_dtype=dtype)
    torch.cuda.synchronize()
    print(model.device_model.kernel_name_for(12 * args.nside**2, args.precision))


if __name__ == "__main__":
    main()
