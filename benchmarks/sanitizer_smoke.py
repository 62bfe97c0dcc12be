"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck).

    compute-sanitizer --tool memcheck python benchmarks/sanitizer_smoke.py
"""
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import zodipy_b200 as zp  # noqa: E402
from zodipy_b200 import engine  # noqa: E402

EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])
Q = zp.Quantity


def main() -> None:
    rng = np.random.default_rng(0)
    for n in (1, 777, 2 * 148 * 512 + 13):  # tiny (lane-split kernels), ragged, large enough for the packed kernel
        u = rng.normal(size=(3, n))
        u /= np.linalg.norm(u, axis=0)
        for name, x, unit in (("dirbe", 25.0, "um"), ("dirbe", 1.25, "um"), ("planck18", 857.0, "GHz"),
                              ("rrm-experimental", 25.0, "um")):
            for precision in ("fp32", "fp64"):
                m = zp.Model(Q(x, unit), name=name, precision=precision)
                m.evaluate_xyz(u, EARTH, return_comps=True)
                m.evaluate_xyz(torch.as_tensor(u, device="cuda"), torch.as_tensor(EARTH, device="cuda"))
    m = zp.Model(Q(25.0, "um"), precision="fp32")
    m.evaluate_healpix(64, EARTH, out_dtype=np.float32)
    m.evaluate_healpix(16, EARTH, nest=True, precision="fp64")
    mb = zp.MultiBandModel([Q(x, "um") for x in (1.25, 12.0, 25.0, 60.0)], precision="fp64")
    mb.evaluate_xyz(u[:, :5000], EARTH)
    mb32 = zp.MultiBandModel([Q(x, "GHz") for x in (353.0, 545.0, 857.0)], name="planck18", precision="fp32")
    mb32.evaluate_xyz(u[:, :5000], EARTH)
    g = np.linspace(-3.0, 3.0, 24)
    zp.grid_number_density_xyz(g, g, g[:12] * 0.2, EARTH[:, 0], model="rrm-experimental")
    zp.grid_number_density_xyz(g, g, g[:12] * 0.2, EARTH[:, 0], model="dirbe")
    t0, dt = 59215.0, 1.0 / 24.0
    tk = t0 + dt * np.arange(24 * 20 + 1)
    lon = 2 * np.pi * (tk - t0) / 365.25 + 1.7
    knots = np.array([np.cos(lon), np.sin(lon), 1e-5 * np.sin(3 * lon)])
    eph = engine.DeviceEphemeris(t0, dt, knots)
    t = np.sort(rng.uniform(t0, tk[-1], 5000))
    m.evaluate_tod_xyz(np.ascontiguousarray(u[:, :5000]), t, eph, observer="semb-l2")
    torch.cuda.synchronize()
    print("sanitizer smoke done")


if __name__ == "__main__":
    main()
