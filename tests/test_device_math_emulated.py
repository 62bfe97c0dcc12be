"""CPU check of the arithmetic the CUDA kernels run.

tests/host_emu/zodi_emu.cpp compiles the SAME header the kernels are built from
(zodipy_b200/csrc/zodi_device.cuh + zodi_model_build.hpp) for the host, so the descriptor ->
device-constant derivation and the fused per-line-of-sight routine are compared with the
reference's outputs without a GPU.  This is a test tool: the product never loads it.
The real parity gate is tests/test_gpu_parity.py (-m gpu) through the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import zodi_oracle as oracle
from helpers import COMP_FLOOR_FP64, TOL_FP32, TOL_FP64, case_ids, golden_case, max_rel_comps, max_rel_total
from zodipy_b200.spec import pack_desc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "zodi_emu.cpp")
LIB = os.path.join(HERE, "host_emu", "libzodi_emu.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "zodipy_b200", "csrc", f)
                for f in ("zodi_device.cuh", "zodi_model_build.hpp", "zodi_kelsall.cuh", "zodi_kelsall_x2.cuh",
                          "zodi_rrm.cuh", "zodi_rrm_x2.cuh", "zodi_multiband.cuh", "zodi_multiband_x2.cuh")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC],
                       check=True)
    return C.CDLL(LIB)


def run_emu(emu, spec, u, obs, earth, precision, lanes, fast=None):
    """fast=None: generic routine; 1: scalar fused routine; 2: packed fused routines (when eligible).
    Returns (emission, which routine ran)."""
    if fast is not None:
        return _run_emu_mode(emu, spec, u, obs, earth, precision, lanes, fast)
    desc, keep = pack_desc(spec)
    u, obs, earth = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, obs, earth))
    flags = oracle.outside_flags(spec, obs)
    out = np.zeros((len(spec["comps"]), u.shape[1]))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = emu.zodi_emu_evaluate(C.byref(desc), precision, lanes, C.c_int64(u.shape[1]), ptr(u), ptr(obs),
                               C.c_int64(obs.shape[1]), ptr(earth), C.c_int64(earth.shape[1]),
                               ptr(flags), ptr(out))
    assert rc == 0
    return out


def _run_emu_mode(emu, spec, u, obs, earth, precision, lanes, fast):
    desc, keep = pack_desc(spec)
    u, obs, earth = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, obs, earth))
    flags = oracle.outside_flags(spec, obs)
    out = np.zeros((len(spec["comps"]), u.shape[1]))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    used = emu.zodi_emu_evaluate_mode(C.byref(desc), precision, lanes, fast, C.c_int64(u.shape[1]), ptr(u),
                                      ptr(obs), C.c_int64(obs.shape[1]), ptr(earth),
                                      C.c_int64(earth.shape[1]), ptr(flags), ptr(out))
    return out, used


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("case_id", case_ids())
def test_fused_routines_match_reference(emu, case_id, precision):
    """Scalar fused Kelsall routine (zodi_kelsall.cuh), fp64 and fp32."""
    case, a = golden_case(case_id)
    em, used = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], precision, 1, fast=1)
    if not used:
        pytest.skip("model layout takes the generic routine")
    tol, floor = (TOL_FP64, COMP_FLOOR_FP64) if precision == 0 else (TOL_FP32, 1.0)
    assert max_rel_total(em, a["emission"]) <= tol
    assert max_rel_comps(em, a["emission"], floor=floor) <= tol


@pytest.mark.parametrize("lanes", [1, 2, 8])
@pytest.mark.parametrize("case_id", case_ids())
def test_packed_routines_equal_scalar_fused(emu, case_id, lanes):
    """Packed routines (zodi_kelsall_x2.cuh; cloud + bands, ring | ring, feature | feature, any lane
    split) perform the same operations as the scalar fused ones."""
    case, a = golden_case(case_id)
    packed, used = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, lanes, fast=2)
    if used != 2:
        pytest.skip("not eligible for the packed routines (generic layout)")
    scalar, _ = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, lanes, fast=1)
    # bit-identical on the GPU (ex2.approx.ftz flushes to 0); the host's libm returns denormals
    # where one lane of a pair is beyond the underflow threshold, hence the 1e-37 allowance
    np.testing.assert_allclose(packed, scalar, rtol=0, atol=1e-37)


@pytest.mark.parametrize("lanes", [1, 8])
@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("case_id", [c for c in case_ids() if "rrm" in c])
def test_fused_rrm_routine_matches_reference(emu, case_id, precision, lanes):
    """Fused RRM routine (zodi_rrm.cuh: grouped bands, shared log2 R^2) against the reference's outputs."""
    case, a = golden_case(case_id)
    em, used = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], precision, lanes, fast=1)
    assert used == 3, "the shipped RRM layout must take the fused routine"
    tol, floor = (TOL_FP64, COMP_FLOOR_FP64) if precision == 0 else (TOL_FP32, 1.0)
    assert max_rel_total(em, a["emission"]) <= tol
    assert max_rel_comps(em, a["emission"], floor=floor) <= tol


@pytest.mark.parametrize("case_id", [c for c in case_ids() if "rrm" in c])
def test_packed_rrm_routines_match_scalar_and_reference(emu, case_id):
    """Packed RRM routines (zodi_rrm_x2.cuh) against the scalar fused routine (same operations per line of
    sight) and the reference's outputs."""
    case, a = golden_case(case_id)
    packed, used = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, 1, fast=2)
    assert used == 4
    scalar, _ = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, 1, fast=1)
    np.testing.assert_allclose(packed, scalar, rtol=2e-6, atol=1e-30)
    assert max_rel_total(packed, a["emission"]) <= TOL_FP32
    assert max_rel_comps(packed, a["emission"], floor=1.0) <= TOL_FP32


@pytest.mark.parametrize("case_id", case_ids())
def test_fp64_arithmetic_matches_reference(emu, case_id):
    case, a = golden_case(case_id)
    em = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 0, 1)
    assert max_rel_total(em, a["emission"]) <= TOL_FP64
    assert max_rel_comps(em, a["emission"], floor=COMP_FLOOR_FP64) <= TOL_FP64


@pytest.mark.parametrize("lanes", [2, 8, 32])
def test_fp64_lane_split_is_equivalent(emu, lanes):
    case, a = golden_case("dirbe_25um_rand")
    em = run_emu(emu, case["spec"], a["u"][:, :200], a["obs"], a["earth"], 0, lanes)
    assert max_rel_total(em, a["emission"][:, :200]) <= TOL_FP64


@pytest.mark.parametrize("case_id", case_ids())
def test_fp32_arithmetic_within_tolerance(emu, case_id):
    """fp32 formulation (libm instead of MUFU): total within 1e-5; components within 1e-5 of
    max(|component|, |total|) (SURVEY 8(a) fp32 note)."""
    case, a = golden_case(case_id)
    em = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, 1)
    assert max_rel_total(em, a["emission"]) <= TOL_FP32
    assert max_rel_comps(em, a["emission"], floor=1.0) <= TOL_FP32


@pytest.mark.parametrize("n", [4, 5, 25, 200])
def test_spline_builder_matches_scipy(emu, n):
    """The library's not-a-knot spline (ephemeris interpolation, zodipy/bodies.py:29-35) against
    scipy.interpolate.CubicSpline, uniform and non-uniform knots."""
    from scipy.interpolate import CubicSpline

    rng = np.random.default_rng(n)
    for x in (59215.0 + np.arange(n) / 24.0, np.sort(rng.uniform(0, 10, n))):
        y = np.cos(0.7 * (x - x[0])) + 0.1 * rng.standard_normal(n)
        c = np.zeros((4, n - 1))
        x_c, y_c = np.ascontiguousarray(x), np.ascontiguousarray(y)
        emu.zodi_emu_spline(n, x_c.ctypes.data_as(C.c_void_p), y_c.ctypes.data_as(C.c_void_p),
                            c.ctypes.data_as(C.c_void_p))
        ref = CubicSpline(x, y).c
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert np.max(np.abs(c - ref) / scale) < 1e-11


def _math(emu, op, x, aux=0.0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert emu.zodi_emu_math(C.c_int(op), C.c_int64(x.size), ptr(x), C.c_double(aux), ptr(y)) == 0
    return y


def test_fp64_log2_table_routine(emu):
    """Math<double>::log2_ (512-bin table + degree-3 polynomial): ~2 ulp over 120 octaves."""
    rng = np.random.default_rng(5)
    x = np.exp2(rng.uniform(-60, 60, 200_000))
    err = np.abs(_math(emu, 0, x) - np.log2(x).astype(np.longdouble).astype(np.float64))
    assert err.max() <= 4e-16 * np.maximum(1.0, np.abs(np.log2(x))).max()
    near_one = 1.0 + rng.uniform(-1e-3, 1e-3, 100_000)  # |log2| small: absolute error matters
    assert np.abs(_math(emu, 0, near_one) - np.log2(near_one)).max() <= 3e-16
    special = _math(emu, 0, np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1.0, 2.0, 0.5]))
    assert special[0] == -np.inf and np.isnan(special[1]) and special[2] == np.inf and np.isnan(special[3])
    assert special[4] == -1074.0  # denormal: library path
    np.testing.assert_allclose(special[5:], [0.0, 1.0, -1.0], rtol=0, atol=2.5e-16)  # table path, not exact


def test_fp64_exp2_table_routine(emu):
    """Math<double>::exp2_ (1024-bin table + degree-2 polynomial): <= 5e-16 relative on (-1020, 1020),
    exact zero below, NaN propagated, exact powers of two exact."""
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-1019.9, 1019.9, 200_000), rng.uniform(-3, 3, 200_000)])
    got, ref = _math(emu, 1, x), np.exp2(x)
    assert (np.abs(got - ref) / ref).max() <= 5e-16
    k = np.arange(-1019, 1020, dtype=np.float64)
    np.testing.assert_array_equal(_math(emu, 1, k), np.exp2(k))
    low = _math(emu, 1, np.array([-1020.0, -1020.5, -1075.0, -1e9, -1e300, -np.inf]))
    np.testing.assert_array_equal(low, 0.0)
    assert np.isnan(_math(emu, 1, np.array([np.nan, -np.nan]))).all()


def test_fp64_table_coordinate_clamps(emu):
    """table_coord<double>: np.interp's clamping expressed on the integer index."""
    top = 99.0
    t = np.array([-5.0, -1e-9, 0.0, 0.25, 41.75, 98.999, 99.0, 99.5, 1e12, np.inf])
    idx, frac = _math(emu, 4, t, top), _math(emu, 3, t, top)
    np.testing.assert_array_equal(idx, [0, 0, 0, 0, 41, 98, 99, 99, 99, 99])
    np.testing.assert_allclose(frac[:6], [0.0, 0.0, 0.0, 0.25, 0.75, 0.999], rtol=0, atol=1e-12)
    assert np.isnan(_math(emu, 3, np.array([np.nan]), top)[0])  # NaN temperature stays NaN


def test_fp64_atan2_abs_routine(emu):
    """Math<double>::atan2_abs_ (65-entry table + series): |atan2(y, x)| to ~2 ulp of pi, all octants."""
    rng = np.random.default_rng(8)
    for x in (1.0, -1.0, 0.37, -2.5, 1e-3, -1e-3, 0.0):
        y = np.concatenate([rng.uniform(-3, 3, 50_000), rng.uniform(-1e-3, 1e-3, 10_000),
                            [0.0, x, -x, 1e-300, 5.0]])
        got = _math(emu, 5, y, aux=x)
        ref = np.abs(np.arctan2(y, x))
        assert np.abs(got - ref).max() <= 9e-16, x
        small = np.abs(y) < 1e-2 * abs(x)
        if x > 0 and small.any():  # near the axis the angle itself is small: relative accuracy
            assert (np.abs(got[small] - ref[small]) <= 4e-16 * ref[small] + 1e-300).all()
    assert _math(emu, 5, np.array([0.0]), aux=0.0)[0] == 0.0


def test_fp32_asin_routine(emu):
    c = np.concatenate([np.linspace(-1, 1, 200_001), [1e-30, -1e-30, 0.5, -0.5, 0.50000006, 0.0]])
    got = _math(emu, 2, c)
    ref = np.arcsin(c.astype(np.float32).astype(np.float64))
    assert np.abs(got - ref).max() <= 2.5e-7  # ~2 ulp of pi/2
    small = np.abs(c) < 0.5
    assert (np.abs(got[small] - ref[small]) <= 1.3e-7 * np.abs(ref[small]) + 1e-45).all()  # relative near 0


@pytest.mark.parametrize("coeffs", [(-0.942, 0.121, -0.165), (-0.527, 0.187, -0.598), (-0.431, 0.172, -0.633),
                                    (0.3, 0.05, -3.0)])
def test_fp32_phase_function_polynomial(emu, coeffs):
    """phase_of_cos<float>: C1 + C2 Theta + exp(C3 Theta) at Theta = arccos(-c) without cancellation.
    The first three coefficient sets are DIRBE 1.25 / 2.2 / 3.5 um (source_params.py), the last one
    forces the literal fallback (|C3| too large for 14 Taylor terms)."""
    C1, C2, C3 = coeffs
    c = np.linspace(-1, 1, 20_001)
    y = np.empty_like(c)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    terms = emu.zodi_emu_phase(C.c_double(C1), C.c_double(C2), C.c_double(C3), C.c_int64(c.size), ptr(c), ptr(y))
    theta = np.arccos(-c.astype(np.float32).astype(np.float64))
    ref = C1 + C2 * theta + np.exp(C3 * theta)
    scale = np.abs(C1) + np.abs(C2) * np.pi + 1.0
    if abs(C3) < 1.0:
        assert terms in (8, 14)
        assert (np.abs(y - ref) / np.abs(ref)).max() <= 2e-6  # relative to the (cancelled) value itself
    else:
        assert terms == 0
        assert (np.abs(y - ref) / scale).max() <= 1e-6  # literal form: relative to the terms' size


MULTIBAND_SETS = [
    ("dirbe", [1.25, 2.2, 3.5, 4.9, 12.0, 25.0, 60.0, 100.0, 140.0, 240.0], "um"),  # scattering in 3 of 10 bands
    ("planck18", [100.0, 143.0, 217.0, 353.0, 545.0, 857.0], "GHz"),
    ("dirbe", [25.0, 60.0, 100.0], "um"),
    ("planck13", [353.0, 545.0, 857.0], "GHz"),            # four components, other band geometry
    ("odegard", [100.0, 143.0, 217.0, 300.0, 353.0, 450.0, 545.0, 700.0, 857.0], "GHz"),  # NB = 16, 9 bands
]


def run_emu_multiband(emu, specs, u, obs, earth, precision, packed):
    import zodipy_b200._cabi as cabi

    descs = (cabi.ModelDesc * len(specs))()
    keep = []
    for i, sp in enumerate(specs):
        descs[i], k = pack_desc(sp)
        keep.append(k)
    u, obs, earth = (np.ascontiguousarray(a, dtype=np.float64) for a in (u, obs, earth))
    flags = oracle.outside_flags(specs[0], obs)
    out = np.zeros((len(specs), u.shape[1]))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nb = emu.zodi_emu_multiband(descs, len(specs), precision, packed, C.c_int64(u.shape[1]), ptr(u), ptr(obs),
                                C.c_int64(obs.shape[1]), ptr(earth), C.c_int64(earth.shape[1]), ptr(flags), ptr(out))
    assert nb in (4, 8, 16)
    return out


@pytest.mark.parametrize("name,xs,unit", MULTIBAND_SETS)
@pytest.mark.parametrize("mode", ["fp64", "fp32", "fp32-packed"])
def test_multiband_routines_match_oracle(emu, name, xs, unit, mode):
    """Multi-band routines (zodi_multiband.cuh, and the packed form zodi_multiband_x2.cuh with its knot-major
    table rows, per-band scattering mask and warp-skipped band densities) against the oracle, band by band;
    an odd number of lines of sight exercises the packed routine's padded last pair."""
    import zodipy_b200 as zp

    rng = np.random.default_rng(5)
    u = rng.normal(size=(3, 301))
    u /= np.linalg.norm(u, axis=0)
    obs = np.array([[-0.3919640703], [0.9020953332], [0.0005]])
    mb = zp.MultiBandModel([zp.Quantity(x, unit) for x in xs], name=name)
    precision, packed = {"fp64": (0, 0), "fp32": (1, 0), "fp32-packed": (1, 1)}[mode]
    got = run_emu_multiband(emu, mb.specs, u, obs, obs, precision, packed)
    tol = TOL_FP64 if precision == 0 else TOL_FP32
    for b, sp in enumerate(mb.specs):
        ref = oracle.evaluate(sp, u, obs, obs).sum(axis=0)
        assert np.max(np.abs(got[b] - ref) / np.abs(ref)) <= tol, (xs[b], mode)


def test_multiband_packed_time_ordered_positions(emu):
    """Packed multi-band routine with per-sample observer / Earth positions (ring and feature follow the Earth)."""
    import zodipy_b200 as zp

    rng = np.random.default_rng(6)
    n = 64
    u = rng.normal(size=(3, n))
    u /= np.linalg.norm(u, axis=0)
    ang = np.linspace(0.0, 2 * np.pi, n, endpoint=False)
    earth = np.stack([np.cos(ang), np.sin(ang), 0.001 * np.sin(3 * ang)])
    obs = earth * 1.01
    mb = zp.MultiBandModel([zp.Quantity(x, "um") for x in (3.5, 25.0, 140.0)], name="dirbe")
    got = run_emu_multiband(emu, mb.specs, u, obs, earth, 1, 1)
    for b, sp in enumerate(mb.specs):
        ref = oracle.evaluate(sp, u, obs, earth).sum(axis=0)
        assert np.max(np.abs(got[b] - ref) / np.abs(ref)) <= TOL_FP32
