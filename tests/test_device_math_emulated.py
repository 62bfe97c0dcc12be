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
                for f in ("zodi_device.cuh", "zodi_model_build.hpp")]


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


@pytest.mark.parametrize("case_id", case_ids())
def test_packed_routines_equal_scalar_fused(emu, case_id):
    """Packed routines (zodi_kelsall_x2.cuh) perform the same operations as the scalar fused ones."""
    case, a = golden_case(case_id)
    packed, used = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, 1, fast=2)
    if used != 2:
        pytest.skip("not eligible for the packed routines (generic layout or scattering)")
    scalar, _ = run_emu(emu, case["spec"], a["u"], a["obs"], a["earth"], 1, 1, fast=1)
    # bit-identical on the GPU (ex2.approx.ftz flushes to 0); the host's libm returns denormals
    # where one lane of a pair is beyond the underflow threshold, hence the 1e-37 allowance
    np.testing.assert_allclose(packed, scalar, rtol=0, atol=1e-37)


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
