"""Randomised parity sweep: CUDA path vs oracle for random models, wavelengths, observers,
quadrature degrees and pointing sets (seeded; complements the fixed fixtures).  The reference's own
property tests (tests/test_evaluate.py:69-95) only assert shapes; here the values are compared."""
import os

import numpy as np
import pytest

import zodi_oracle as oracle
import zodipy_b200 as zp
from helpers import COMP_FLOOR_FP32, COMP_FLOOR_FP64, TOL_FP32, TOL_FP64, max_rel_total

pytestmark = pytest.mark.gpu

MODELS = {"dirbe": ("um", 1.25, 240.0), "planck13": ("GHz", 100.0, 857.0), "planck15": ("GHz", 100.0, 857.0),
          "planck18": ("GHz", 100.0, 857.0), "odegard": ("GHz", 100.0, 857.0),
          "rrm-experimental": ("um", 12.0, 100.0)}


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    name = list(MODELS)[seed % len(MODELS)]
    unit, lo, hi = MODELS[name]
    if rng.random() < 0.3:  # bandpass
        centre = np.exp(rng.uniform(np.log(lo * 1.2), np.log(hi / 1.2)))
        x = np.linspace(centre / 1.15, centre * 1.15, int(rng.integers(5, 30)))
        x = x[(x >= lo) & (x <= hi)]
        w = np.exp(-0.5 * ((x - centre) / (0.08 * centre)) ** 2) + 0.05
        model_args = dict(x=zp.Quantity(x, unit), weights=w)
    else:
        model_args = dict(x=zp.Quantity(float(np.exp(rng.uniform(np.log(lo), np.log(hi)))), unit))
    deg = int(rng.choice([5, 11, 32, 50, 50, 50, 64, 101, 150]))
    n = int(rng.integers(200, 3000))
    u = rng.normal(size=(3, n))
    u /= np.linalg.norm(u, axis=0)
    r_obs = float(np.exp(rng.uniform(np.log(0.3), np.log(4.0))))
    lon = rng.uniform(0, 2 * np.pi)
    base = r_obs * np.array([np.cos(lon), np.sin(lon), rng.uniform(-0.05, 0.05)])
    earth0 = np.array([np.cos(lon + 0.3), np.sin(lon + 0.3), 0.0])
    if rng.random() < 0.4:  # time-ordered: per-sample observer / Earth
        ang = np.linspace(0, rng.uniform(0.05, 1.5), n)
        rot = lambda v: np.array([v[0] * np.cos(ang) - v[1] * np.sin(ang), v[0] * np.sin(ang) + v[1] * np.cos(ang),
                                  v[2] + 0 * ang])  # noqa: E731
        obs, earth = rot(base) * (1 + 0.01 * np.sin(7 * ang)), rot(earth0)
    else:
        obs, earth = base.reshape(3, 1), earth0.reshape(3, 1)
    return name, model_args, deg, u, np.ascontiguousarray(obs), np.ascontiguousarray(earth)


def _near_hard_cutoff(spec, u, obs, margin=1e-6):
    """Lines of sight with a quadrature node within `margin` AU of a hard density cut-off
    (R_inner / R_outer of the RRM fan, comet and bands, number_density.py:201-203,239-241,...).
    The reference decides `R <= R_outer` in float64; an fp32 kernel cannot classify a node that
    close to the discontinuity the same way, so those rays are excluded from the fp32 comparison.
    They only occur because the reference's range formula (no z-term, quirk Q2) lets rays of
    observers far off the ecliptic overshoot the cut-off sphere."""
    bad = np.zeros(u.shape[1], dtype=bool)
    start, stop = oracle.los_range(spec, u, obs)
    for ci, comp in enumerate(spec["comps"]):
        bounds = [comp["params"][k] for k in ("R_inner", "R_outer") if k in comp["params"]]
        if not bounds:
            continue
        x0 = np.asarray(comp["params"]["X_0"]).reshape(3, 1)
        for r in spec["points"]:
            R_los = 0.5 * (stop[ci] - start[ci]) * r + 0.5 * (stop[ci] + start[ci])
            R = np.sqrt(((R_los * u + obs - x0) ** 2).sum(axis=0))
            for b in bounds:
                bad |= np.abs(R - b) < margin
    return bad


N_SEEDS = int(os.environ.get("ZODI_SWEEP_SEEDS", "36"))  # raise for a wider one-off hunt


@pytest.mark.parametrize("seed", range(N_SEEDS))
def test_random_configuration(seed):
    name, model_args, deg, u, obs, earth = _case(seed)
    ref = noise = None
    for precision, tol, floor in (("fp64", TOL_FP64, COMP_FLOOR_FP64), ("fp32", TOL_FP32, COMP_FLOOR_FP32)):
        model = zp.Model(name=name, gauss_quad_degree=deg, precision=precision, **model_args)
        if ref is None:
            ref = oracle.evaluate(model.spec, u, obs, earth)
        got = model.evaluate_xyz(u, obs, earth, return_comps=True)
        if precision == "fp32" and model.spec["kind"] == "rrm":
            keep = ~_near_hard_cutoff(model.spec, u, obs)
            assert keep.mean() > 0.9  # Gauss-Legendre nodes crowd towards the ends of the range at high degree
            got, ref_p = got[:, keep], ref[:, keep]
        else:
            ref_p = ref
        info = (seed, name, precision, deg, float(np.linalg.norm(obs[:, 0])))
        # totals relative to sum_c |component|: identical to |total| for the usual all-positive models,
        # but planck13's partly NEGATIVE emissivities (source_params.py:43,47-48) let components cancel,
        # so the total can pass through zero and its own relative error is unbounded
        l1 = np.abs(ref_p).sum(axis=0)
        assert np.nanmax(np.abs(got.sum(axis=0) - ref_p.sum(axis=0)) / l1) <= tol, info
        if name != "planck13":
            assert max_rel_total(got, ref_p) <= tol, info
        if precision == "fp64" and noise is None:
            # observers inside ~0.8 delta_r: the reference's literal 1 - exp(-(R/delta_r)^20) carries
            # cancellation noise far above 1e-10 on the band components; allow that noise bound
            noise = oracle.reference_rounding_noise(model.spec, u, obs, earth)
        allowance = noise if precision == "fp64" else 0.0
        scale = np.maximum(np.abs(ref_p), floor * np.abs(ref_p.sum(axis=0))[None, :])
        assert np.nanmax((np.abs(got - ref_p) - allowance) / scale) <= tol, info
