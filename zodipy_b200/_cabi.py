"""ctypes binding of the C ABI in ``include/zodi_b200.h`` (no torch types cross it).

There is NO fallback: if ``libzodi_b200.so`` has not been built (``python -m zodipy_b200.build``)
or no CUDA device is usable, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ABI_VERSION = 2
MAX_COMPS = 16
MAX_NODES = 1024
MAX_TEMPS = 1024
N_SHAPE = 8
MAX_PEERS = 8
MAX_BANDS = 16
IPC_HANDLE_BYTES = 64

KELSALL, RRM = 0, 1
FP64, FP32 = 0, 1
OUT_F64, OUT_F32 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
PEAK_FP32_FMA, PEAK_FP64_FMA, PEAK_MUFU_EX2, PEAK_HBM_COPY = 0, 1, 2, 3

LIB_NAME = "libzodi_b200.so"
# ZODI_B200_LIB: load another build of the same library (A/B measurements of kernel variants only)
LIB_PATH = os.environ.get("ZODI_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

c_double_p = C.POINTER(C.c_double)
c_uint8_p = C.POINTER(C.c_uint8)


class ComponentDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("reserved", C.c_int32),
        ("x0", C.c_double * 3),
        ("sin_Omega", C.c_double), ("cos_Omega", C.c_double),
        ("sin_i", C.c_double), ("cos_i", C.c_double),
        ("shape", C.c_double * N_SHAPE),
        ("cutoff_inner", C.c_double), ("cutoff_outer", C.c_double),
        ("emissivity", C.c_double), ("albedo", C.c_double),
        ("T_0", C.c_double), ("delta", C.c_double),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("kind", C.c_int32), ("n_comps", C.c_int32),
        ("n_nodes", C.c_int32), ("n_temps", C.c_int32), ("reserved", C.c_int32),
        ("T_0", C.c_double), ("delta", C.c_double), ("C1", C.c_double), ("C2", C.c_double),
        ("C3", C.c_double), ("solar_irradiance", C.c_double), ("calibration", C.c_double),
        ("temps", c_double_p), ("bnu", c_double_p), ("nodes", c_double_p), ("weights", c_double_p),
        ("comps", ComponentDesc * MAX_COMPS),
    ]


class EvalArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("u", C.c_void_p), ("u_stride", C.c_int64),
        ("obs", C.c_void_p), ("n_obs", C.c_int64), ("obs_stride", C.c_int64),
        ("earth", C.c_void_p), ("n_earth", C.c_int64), ("earth_stride", C.c_int64),
        ("outside_flags", c_uint8_p),
        ("return_comps", C.c_int32), ("precision", C.c_int32),
        ("out_dtype", C.c_int32), ("memory", C.c_int32),
        ("out", C.c_void_p), ("out_stride", C.c_int64),
        ("stream", C.c_void_p),
        ("n_peers", C.c_int32), ("reserved", C.c_int32),
        ("peer_out", C.c_void_p * MAX_PEERS),
        ("peer_offset", C.c_int64), ("peer_stride", C.c_int64),
        ("cyclic_block", C.c_int64), ("cyclic_parts", C.c_int32), ("cyclic_rank", C.c_int32),
        ("ephemeris", C.c_void_p), ("obstime", C.c_void_p),
    ]


class EphemerisDesc(C.Structure):
    _fields_ = [
        ("n_knots", C.c_int64), ("t0", C.c_double), ("dt", C.c_double),
        ("earth_knots", c_double_p), ("obs_knots", c_double_p), ("obs_scale", C.c_double),
    ]


class HealpixArgs(C.Structure):
    _fields_ = [
        ("base", EvalArgs),
        ("nside", C.c_int64), ("ipix_start", C.c_int64),
        ("nest", C.c_int32), ("has_rot", C.c_int32),
        ("rot", C.c_double * 9),
    ]


class LonLatArgs(C.Structure):
    _fields_ = [
        ("base", EvalArgs),
        ("lon", C.c_void_p), ("lat", C.c_void_p),
        ("has_rot", C.c_int32), ("reserved", C.c_int32),
        ("rot", C.c_double * 9),
    ]


# every symbol include/zodi_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "zodi_abi_version": (C.c_int, []),
    "zodi_last_error": (C.c_char_p, []),
    "zodi_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "zodi_model_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "zodi_model_update": (C.c_int, [C.c_void_p, C.POINTER(ModelDesc)]),
    "zodi_model_destroy": (C.c_int, [C.c_void_p]),
    "zodi_model_kernel_name": (C.c_char_p, [C.c_void_p]),
    "zodi_model_kernel_for": (C.c_char_p, [C.c_void_p, C.c_int64, C.c_int32]),
    "zodi_evaluate": (C.c_int, [C.c_void_p, C.POINTER(EvalArgs)]),
    "zodi_ephemeris_create": (C.c_int, [C.c_int, C.POINTER(EphemerisDesc), C.POINTER(C.c_void_p)]),
    "zodi_ephemeris_set_obs_scale": (C.c_int, [C.c_void_p, C.c_double]),
    "zodi_ephemeris_destroy": (C.c_int, [C.c_void_p]),
    "zodi_ephemeris_coefficients": (C.c_int, [C.c_void_p, c_double_p]),
    "zodi_ephemeris_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "zodi_ephemeris_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, c_double_p]),
    "zodi_ephemeris_release_times": (C.c_int, [C.c_void_p]),
    "zodi_device_math": (C.c_int, [C.c_int, C.c_int32, C.c_int64, c_double_p, C.c_double, c_double_p]),
    "zodi_evaluate_healpix": (C.c_int, [C.c_void_p, C.POINTER(HealpixArgs)]),
    "zodi_healpix_vectors": (C.c_int, [C.c_int, C.c_int64, C.c_int32, C.c_int64, C.c_int64, c_double_p, C.c_void_p,
                                       C.c_int64, C.c_int32, C.c_void_p]),
    "zodi_evaluate_lonlat": (C.c_int, [C.c_void_p, C.POINTER(LonLatArgs)]),
    "zodi_lonlat_vectors": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_double_p, C.c_void_p, C.c_int64,
                                      C.c_int32, C.c_void_p]),
    "zodi_multiband_evaluate_lonlat": (C.c_int, [C.c_void_p, C.POINTER(LonLatArgs)]),
    "zodi_max_observer_radius": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                           C.c_void_p, c_double_p]),
    "zodi_flags_from_radius": (C.c_int, [C.c_void_p, C.c_double, c_uint8_p]),
    "zodi_multiband_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int32, C.c_int, C.POINTER(C.c_void_p)]),
    "zodi_multiband_evaluate": (C.c_int, [C.c_void_p, C.POINTER(EvalArgs)]),
    "zodi_multiband_evaluate_healpix": (C.c_int, [C.c_void_p, C.POINTER(HealpixArgs)]),
    "zodi_multiband_destroy": (C.c_int, [C.c_void_p]),
    "zodi_number_density": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, c_double_p, C.c_void_p,
                                      C.c_int64, C.c_int32, C.c_void_p]),
    "zodi_peer_buffer_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), c_uint8_p]),
    "zodi_peer_buffer_open": (C.c_int, [C.c_int, c_uint8_p, C.POINTER(C.c_void_p)]),
    "zodi_peer_buffer_close": (C.c_int, [C.c_int, C.c_void_p]),
    "zodi_peer_buffer_free": (C.c_int, [C.c_int, C.c_void_p]),
    "zodi_peer_rendezvous": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]),
    "zodi_peak_probe": (C.c_int, [C.c_int, C.c_int32, c_double_p]),
    "zodi_kernel_launch_count": (C.c_int64, []),
    "zodi_last_kernel_ms": (C.c_double, [C.c_void_p]),
}

_lib = None


class ZodiError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def load():
    """Load the CUDA library (once).  Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -m zodipy_b200.build` "
            "(needs nvcc). zodipy_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.zodi_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{LIB_NAME} ABI version {lib.zodi_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().zodi_last_error()
        raise ZodiError(f"zodi_b200 error {status}: {msg.decode() if msg else '?'}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().zodi_device_count(C.byref(n)))
    return n.value


def as_double_p(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)
