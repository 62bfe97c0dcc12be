"""Time the reference's own hot-path modules against the oracle port on identical inputs
(build container only; TEST INFRASTRUCTURE).  Run: python oracle/compare_speed_with_reference.py"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import ctypes  # noqa: E402

if "--tuned-malloc" in sys.argv:  # what bench.py's CPU legs do (tune_malloc_for_numpy)
    _libc = ctypes.CDLL("libc.so.6")
    _libc.mallopt(-3, 1 << 30), _libc.mallopt(-1, 1 << 31), _libc.mallopt(-2, 1 << 28)
import make_golden as mg  # noqa: E402
import zodi_oracle as oracle  # noqa: E402
from zodipy_b200 import healpix  # noqa: E402

ns = mg.load_reference_full()
u = healpix.pix2vec_ring(64, np.arange(12 * 64 * 64))
obs = np.array([[-0.3919640703], [0.9020953332], [0.0]])
for name, x, unit in (("planck18", 857.0, "GHz"), ("dirbe", 25.0, "micron")):
    spec, model, comp_params, shared, comps = mg.build_spec(ns, name, x, unit)
    best_ref = best_port = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ref, _, _ = mg.reference_emission(ns, model, comps, comp_params, shared, spec["table"], spec["points"],
                                          spec["weights"], u, obs, obs)
        best_ref = min(best_ref, time.perf_counter() - t0)
        t0 = time.perf_counter()
        port = oracle.evaluate(spec, u, obs, obs)
        best_port = min(best_port, time.perf_counter() - t0)
    units = u.shape[1] * len(spec["comps"]) * 50
    print(f"{name}: reference modules {best_ref:.3f} s ({units / best_ref:.3g} evals/s/core), oracle port "
          f"{best_port:.3f} s ({units / best_port:.3g} evals/s/core), port/reference speed = {best_ref / best_port:.2f}, "
          f"max |diff| = {np.max(np.abs(ref - port)):.1e}")
