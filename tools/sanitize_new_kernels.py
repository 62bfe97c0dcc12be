"""Target for `compute-sanitizer --tool memcheck`: one launch of each kernel added late in round 2 (packed
multi-band NB = 4 / 8 / 16 incl. scattering, persistent-tile packed Kelsall kernel) on small inputs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZODI_X2_PERSIST"] = "2"
import zodipy_b200 as zp  # noqa: E402

EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])
rng = np.random.default_rng(0)
u = rng.normal(size=(3, 40001))
u /= np.linalg.norm(u, axis=0)
only_persistent = "--persistent-only" in sys.argv
for name, xs, unit in () if only_persistent else (("dirbe", [1.25, 2.2, 3.5, 4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0], "um"),
                       ("planck18", [100.0, 143.0, 217.0, 353.0, 545.0, 857.0], "GHz"),
                       ("dirbe", [25.0, 60.0, 100.0], "um")):
    mb = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name, precision="fp32")
    out = mb.evaluate_xyz(u, EARTH)
    print(mb.device_model.kernel_name_for(u.shape[1], "fp32"), out.shape, bool(np.isfinite(out).all()))
for name, x, unit in (("planck18", 857.0, "GHz"), ("dirbe", 25.0, "um")):
    m = zp.Model(zp.Quantity(x, unit), name=name, precision="fp32")
    out = m.evaluate_healpix(256, EARTH, out_dtype=np.float32)
    print("persistent tiles", name, out.shape, bool(np.isfinite(out).all()))
