"""Handle-level wrapper over the C ABI: one ``DeviceModel`` = one parameter block on one GPU.

This is the array seam of the reference (``zodipy/model.py:253-279``): ecliptic unit vectors
``(3, N)``, observer and Earth positions ``(3, 1)`` or ``(3, N)`` in, emission ``(ncomps, N)`` or
``(N,)`` out.  Inputs may be NumPy arrays (host memory: the library stages them through a
pipelined H2D / kernel / D2H workspace) or torch CUDA tensors (device memory: one asynchronous
kernel launch on the current torch stream, result stays on the GPU).  torch is only used for
device memory and streams; the compute is the library's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi
from .spec import outside_flags as spec_outside_flags
from .spec import pack_desc

_PRECISIONS = {"fp64": _cabi.FP64, "fp32": _cabi.FP32}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _rows(a, name: str):
    """(data_ptr, row_length, row_stride) of a (3, n) float64 array with unit inner stride."""
    if _is_torch(a):
        import torch

        if a.dtype != torch.float64 or a.dim() != 2 or a.shape[0] != 3:
            raise ValueError(f"{name} must be a (3, n) float64 tensor")
        if a.shape[1] > 1 and a.stride(1) != 1:
            a = a.contiguous()
        return a, a.data_ptr(), a.shape[1], (a.stride(0) if a.shape[1] > 1 else max(a.stride(0), 1))
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    if a.ndim != 2 or a.shape[0] != 3:
        raise ValueError(f"{name} must have shape (3, n)")
    if a.shape[1] > 1 and a.strides[1] != 8:
        a = np.ascontiguousarray(a)
    stride = a.strides[0] // 8 if a.shape[1] > 1 else max(a.strides[0] // 8, 1)
    return a, a.ctypes.data, a.shape[1], stride


def _vec(a, name: str):
    """(array, data_ptr, n) of a contiguous 1-D float64 array (NumPy or torch CUDA)."""
    if _is_torch(a):
        import torch

        if a.dtype != torch.float64:
            raise ValueError(f"{name} must be float64")
        a = a.reshape(-1).contiguous()
        return a, a.data_ptr(), a.shape[0]
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    return a, a.ctypes.data, a.size


class _LonLat:
    """Directions given as spherical coordinates [rad] + optional 3x3 rotation to the mean ecliptic."""

    def __init__(self, lon, lat, rot=None):
        if _is_torch(lon) != _is_torch(lat):
            raise ValueError("lon and lat must both be NumPy arrays or both torch CUDA tensors")
        self.lon, self.lon_ptr, self.n = _vec(lon, "lon")
        self.lat, self.lat_ptr, n_lat = _vec(lat, "lat")
        if n_lat != self.n:
            raise ValueError("lon and lat must have the same length")
        self.rot = None if rot is None else np.ascontiguousarray(np.asarray(rot, dtype=np.float64).reshape(9))

    def pack(self, base):
        ll = _cabi.LonLatArgs()
        ll.base = base
        ll.lon, ll.lat = self.lon_ptr, self.lat_ptr
        if self.rot is not None:
            ll.has_rot = 1
            for i in range(9):
                ll.rot[i] = float(self.rot[i])
        return ll


MEAN_DIST_TO_L2 = 0.009896235034000056  # AU, zodipy/bodies.py:13


class DeviceEphemeris:
    """Earth / observer ephemeris as device-resident cubic splines (time-ordered data).

    ``earth_knots`` (3, K): Earth's heliocentric ecliptic position [AU] at the uniformly spaced times
    ``t0 + k * dt`` - what the reference samples hourly with Astropy and interpolates with SciPy's
    ``CubicSpline`` (``zodipy/bodies.py:16-35``).  ``obs_knots``: the observer's knots, or ``None``
    for an observer tied to Earth: ``obspos="earth"`` (scale 1) or ``"semb-l2"`` (scale set per
    evaluation from the samples, reproducing ``bodies.py:38-50`` including its un-axised norm).
    """

    def __init__(self, t0: float, dt: float, earth_knots, obs_knots=None, device: int = 0):
        self._lib = _cabi.load()
        self.device = int(device)
        earth = np.ascontiguousarray(earth_knots, dtype=np.float64)
        if earth.ndim != 2 or earth.shape[0] != 3:
            raise ValueError("earth_knots must have shape (3, n_knots)")
        d = _cabi.EphemerisDesc()
        d.n_knots, d.t0, d.dt, d.obs_scale = earth.shape[1], float(t0), float(dt), 1.0
        d.earth_knots = _cabi.as_double_p(earth)
        obs = None
        if obs_knots is not None:
            obs = np.ascontiguousarray(obs_knots, dtype=np.float64)
            if obs.shape != earth.shape:
                raise ValueError("obs_knots must have the shape of earth_knots")
            d.obs_knots = _cabi.as_double_p(obs)
        self.n_knots, self.t0, self.dt = earth.shape[1], float(t0), float(dt)
        self.has_obs_knots = obs is not None
        self._handle = C.c_void_p()
        _cabi.check(self._lib.zodi_ephemeris_create(self.device, C.byref(d), C.byref(self._handle)))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.zodi_ephemeris_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_obs_scale(self, scale: float) -> None:
        _cabi.check(self._lib.zodi_ephemeris_set_obs_scale(self._handle, float(scale)))

    def coefficients(self) -> np.ndarray:
        """Earth spline coefficients in scipy's ``CubicSpline.c`` layout (4, n_knots - 1, 3)."""
        c = np.empty((4, self.n_knots - 1, 3), dtype=np.float64)
        _cabi.check(self._lib.zodi_ephemeris_coefficients(self._handle, _cabi.as_double_p(c)))
        return c

    def _times(self, t):
        if _is_torch(t):
            import torch

            if t.dtype != torch.float64 or not t.is_contiguous() or not t.is_cuda:
                raise ValueError("device obstime must be a contiguous float64 CUDA tensor")
            return t, t.data_ptr(), t.numel(), _cabi.MEM_DEVICE, torch.cuda.current_stream(t.device).cuda_stream
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1)
        return t, t.ctypes.data, t.size, _cabi.MEM_HOST, None

    def positions(self, t):
        """(earth, obs) positions (3, n) at times ``t`` from the device splines (NumPy in/out)."""
        t_a, ptr, n, mem, stream = self._times(np.asarray(t, dtype=np.float64))
        earth, obs = np.empty((3, n)), np.empty((3, n))
        _cabi.check(self._lib.zodi_ephemeris_positions(self._handle, ptr, n, mem, earth.ctypes.data,
                                                       obs.ctypes.data, stream))
        return earth, obs

    def stats(self, t):
        """(sum |earth|^2, max |earth|, max |observer|) over the samples at times ``t``."""
        t_a, ptr, n, mem, stream = self._times(t)
        out = (C.c_double * 3)()
        _cabi.check(self._lib.zodi_ephemeris_stats(self._handle, ptr, n, mem, stream, out))
        return out[0], float(np.sqrt(out[1])), float(np.sqrt(out[2]))

    def release_times(self) -> None:
        """Free the device copy of the sample times that ``stats`` keeps for host arrays (8 B per
        sample, reused by later calls; freed with the ephemeris otherwise)."""
        _cabi.check(self._lib.zodi_ephemeris_release_times(self._handle))

    def prepare(self, t, observer: str = "earth"):
        """Set the observer rule for the samples at times ``t`` and return the largest observer
        distance (for the global early-out flags).  observer: "earth" | "semb-l2" | "knots"."""
        if observer == "knots":
            if not self.has_obs_knots:
                raise ValueError("this ephemeris has no observer knots")
            return self.stats(t)[2]
        if self.has_obs_knots:
            raise ValueError("this ephemeris carries observer knots; use observer='knots'")
        if observer not in ("earth", "semb-l2"):
            raise ValueError("observer must be 'earth', 'semb-l2' or 'knots'")
        self.set_obs_scale(1.0)
        sum_r2, max_earth, _ = self.stats(t)
        scale = 1.0
        if observer == "semb-l2":
            # get_semb_l2_pos: earth / ||earth|| * (||earth|| + L2) with the norm over the WHOLE array
            norm = float(np.sqrt(sum_r2))
            scale = (norm + MEAN_DIST_TO_L2) / norm
            self.set_obs_scale(scale)
        return scale * max_earth


class DeviceModel:
    """Device-resident model parameters + the evaluate call."""

    def __init__(self, spec: dict, device: int = 0):
        self._lib = _cabi.load()
        self.spec = spec
        self.device = int(device)
        self.ncomps = len(spec["comps"])          # rows of a return_comps output
        self.n_model_comps = len(spec["comps"])   # components of the model (rows of outside_flags)
        self._handle = C.c_void_p()
        desc, keep = pack_desc(spec)
        _cabi.check(self._lib.zodi_model_create(C.byref(desc), self.device, C.byref(self._handle)))
        del keep

    def update(self, spec: dict) -> None:
        """Re-upload after ``Model.update_parameters`` (``zodipy/model.py:313-333``)."""
        desc, keep = pack_desc(spec)
        _cabi.check(self._lib.zodi_model_update(self._handle, C.byref(desc)))
        self.spec = spec
        self.ncomps = self.n_model_comps = len(spec["comps"])
        del keep

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.zodi_model_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    def max_observer_radius(self, obs) -> float:
        obs_a, ptr, n_obs, stride = _rows(obs, "obs")
        out = C.c_double(0.0)
        if _is_torch(obs_a):
            import torch

            mem, stream = _cabi.MEM_DEVICE, torch.cuda.current_stream(obs_a.device).cuda_stream
        else:
            mem, stream = _cabi.MEM_HOST, None
        _cabi.check(self._lib.zodi_max_observer_radius(self._handle, ptr, n_obs, stride, mem, stream,
                                                       C.byref(out)))
        return out.value

    def outside_flags(self, obs) -> np.ndarray:
        return spec_outside_flags(self.spec, self.max_observer_radius(obs))

    def evaluate(self, u, obs=None, earth=None, *, return_comps: bool = False, precision: str = "fp64",
                 out=None, out_dtype=None, outside_flags=None, peer_map=None, ephemeris=None,
                 obstime=None, observer: str = "earth", lonlat=None):
        """Emission [MJy/sr] for unit vectors ``u`` (3, N).

        ``lonlat``: a :class:`_LonLat` (see :meth:`evaluate_lonlat`); ``u`` is then ``None`` and the
        unit vectors are formed (and rotated to the ecliptic) in the kernel prologue.

        ``obs`` / ``earth``: (3,), (3, 1) or (3, N) [AU]; ``earth`` defaults to ``obs``.
        ``outside_flags``: optional (ncomps, 2) uint8 GLOBAL early-out flags (needed when the
        observers of a job are sharded over several calls/GPUs); default: derived from ``obs``.
        ``ephemeris`` + ``obstime``: time-ordered data with positions evaluated on the device from a
        :class:`DeviceEphemeris` at the per-sample times ``obstime`` (n,) (``obs`` / ``earth`` are
        then not needed); ``observer`` selects "earth", "semb-l2" or the ephemeris' own observer
        "knots".
        ``peer_map``: a :class:`zodipy_b200.sharding.PeerMap`; the kernel then stores this call's
        slice directly into every GPU's full map (fused all-gather) and nothing is returned.
        Returns an array like the inputs (NumPy -> NumPy, torch CUDA -> torch CUDA) of shape
        (ncomps, N) if ``return_comps`` else (N,).
        """
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        if lonlat is not None:
            device_mem = _is_torch(lonlat.lon)
            u_a, u_ptr, n, u_stride = lonlat.lon, None, lonlat.n, 0
        else:
            device_mem = _is_torch(u)
            u_a, u_ptr, n, u_stride = _rows(u, "unit_vectors")
        if ephemeris is not None:
            return self._evaluate_tod(u_a, u_ptr, n, u_stride, device_mem, ephemeris, obstime, observer,
                                      return_comps, precision, out, out_dtype, outside_flags, lonlat)
        if obs is None:
            raise ValueError("obs is required unless an ephemeris is given")
        if earth is None:
            earth = obs
        if device_mem:
            import torch

            if not u_a.is_cuda:
                raise ValueError("torch inputs must be CUDA tensors (NumPy arrays for host memory)")
            if u_a.device.index != self.device:
                raise ValueError(f"inputs are on cuda:{u_a.device.index}, model on cuda:{self.device}")
            as_dev = lambda a: (a if _is_torch(a) else torch.as_tensor(  # noqa: E731
                np.asarray(a, dtype=np.float64).reshape(3, -1), device=u_a.device))
            obs, earth = as_dev(obs), as_dev(earth)
        obs_a, obs_ptr, n_obs, obs_stride = _rows(obs, "obs")
        earth_a, earth_ptr, n_earth, earth_stride = _rows(earth, "earth")
        if n_obs not in (1, n) or n_earth not in (1, n):
            raise ValueError("obs/earth must hold one position or one per unit vector")

        if out_dtype is None:
            out_dtype = np.float64
        out_dtype = np.dtype(out_dtype)
        if out_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("out_dtype must be float64 or float32")
        shape = (self.ncomps, n) if return_comps else (n,)
        if peer_map is not None:
            if not device_mem:
                raise ValueError("peer_map needs device-resident (torch CUDA) inputs")
            if peer_map.dtype != out_dtype or peer_map.rows != (self.ncomps if return_comps else 1):
                raise ValueError("peer_map dtype/rows do not match this call")
            out_ptr, out = None, None
            stream = torch.cuda.current_stream(u_a.device).cuda_stream
        elif device_mem:
            tdtype = torch.float64 if out_dtype == np.float64 else torch.float32
            if out is None:
                out = torch.empty(shape, dtype=tdtype, device=u_a.device)
            elif tuple(out.shape) != shape or out.dtype != tdtype or not out.is_contiguous():
                raise ValueError("out has wrong shape/dtype or is not contiguous")
            out_ptr = out.data_ptr()
            stream = torch.cuda.current_stream(u_a.device).cuda_stream
        else:
            if out is None:
                out = np.empty(shape, dtype=out_dtype)
            elif out.shape != shape or out.dtype != out_dtype or out.strides[-1] != out.itemsize:
                # rows may be strided (a column block of a larger (ncomps, N) array: one shard's output)
                raise ValueError("out has wrong shape/dtype or its last axis is not contiguous")
            out_ptr = out.ctypes.data
            out_row_stride = out.strides[0] // out.itemsize if out.ndim == 2 and n > 1 else n
            stream = None
        if n == 0:
            return out
        if peer_map is not None:
            self._check_peer_slice(peer_map, n)

        if outside_flags is None:
            if device_mem and n_obs == 1:
                # instantaneous observer: avoid a device reduction + sync, read the 3 numbers
                r = float(torch.linalg.vector_norm(obs_a.reshape(3)).item())
                flags = spec_outside_flags(self.spec, r)
            else:
                flags = self.outside_flags(obs_a)
        else:
            flags = self._checked_flags(outside_flags)

        args = _cabi.EvalArgs()
        args.n = n
        args.u, args.u_stride = u_ptr, u_stride
        args.obs, args.n_obs, args.obs_stride = obs_ptr, n_obs, obs_stride
        args.earth, args.n_earth, args.earth_stride = earth_ptr, n_earth, earth_stride
        args.outside_flags = flags.ctypes.data_as(_cabi.c_uint8_p)
        args.return_comps = int(bool(return_comps))
        args.precision = _PRECISIONS[precision]
        args.out_dtype = _cabi.OUT_F64 if out_dtype == np.float64 else _cabi.OUT_F32
        args.memory = _cabi.MEM_DEVICE if device_mem else _cabi.MEM_HOST
        args.out, args.out_stride = out_ptr, (n if device_mem or peer_map is not None else out_row_stride)
        args.stream = stream
        if peer_map is not None:
            args.n_peers = len(peer_map.pointers)
            for i, ptr in enumerate(peer_map.pointers):
                args.peer_out[i] = ptr
            args.peer_offset, args.peer_stride = peer_map.offset, peer_map.n_total
            if peer_map.cyclic is not None:
                args.cyclic_block, args.cyclic_parts, args.cyclic_rank = peer_map.cyclic
        self._dispatch(args, lonlat)
        return out

    def _checked_flags(self, outside_flags) -> np.ndarray:
        """(n_model_comps, 2) uint8 early-out flags: the library reads 2 bytes per MODEL component, whatever
        the number of output rows (a multi-band handle returns one row per band)."""
        flags = np.ascontiguousarray(outside_flags, dtype=np.uint8)
        if flags.shape != (self.n_model_comps, 2):
            raise ValueError(f"outside_flags must have shape ({self.n_model_comps}, 2)")
        return flags

    def _check_peer_slice(self, peer_map, n: int) -> None:
        """The slice this call stores must lie inside every peer's map (the kernel writes remote memory)."""
        if peer_map.cyclic is None:
            if peer_map.offset < 0 or peer_map.offset + n > peer_map.n_total:
                raise ValueError("slice does not fit the peer map")
            return
        from .sharding import cyclic_count

        block, parts, rank = peer_map.cyclic
        if n != cyclic_count(peer_map.n_total - peer_map.offset, parts, rank, block):
            raise ValueError("block-cyclic shard does not match the peer map: this rank must evaluate exactly its "
                             f"share of the {peer_map.n_total - peer_map.offset} mapped lines of sight")

    def _dispatch(self, args, lonlat) -> None:
        if lonlat is None:
            _cabi.check(self._lib.zodi_evaluate(self._handle, C.byref(args)))
        else:
            _cabi.check(self._lib.zodi_evaluate_lonlat(self._handle, C.byref(lonlat.pack(args))))

    def evaluate_lonlat(self, lon, lat, obs=None, earth=None, *, rot=None, **kwargs):
        """Emission for directions given as longitude / latitude [rad] (N,) of a frame whose constant
        rotation to the mean ecliptic is the 3x3 ``rot`` (``None``: the angles are ecliptic).

        The unit vectors ``rot @ (cos lat cos lon, cos lat sin lon, sin lat)`` are formed in the
        kernel prologue, so 16 instead of 24 B per line of sight are uploaded and the host never
        builds the (3, N) array.  All keyword arguments of :meth:`evaluate` apply (per-sample
        ``obs`` / ``earth``, ``ephemeris`` + ``obstime``, ``peer_map``, ...).
        """
        return self.evaluate(None, obs, earth, lonlat=_LonLat(lon, lat, rot), **kwargs)

    def _evaluate_tod(self, u_a, u_ptr, n, u_stride, device_mem, ephemeris, obstime, observer,
                      return_comps, precision, out, out_dtype, outside_flags, lonlat=None):
        """``observer="prepared"``: the caller has already run ``ephemeris.stats`` on these very times
        (which staged them on the device), set the observer scale and supplies ``outside_flags`` - used
        when the samples of one job are split over several devices and the reductions are global."""
        if obstime is None:
            raise ValueError("obstime is required with an ephemeris")
        if ephemeris.device != self.device:
            raise ValueError("ephemeris and model live on different devices")
        if device_mem != _is_torch(obstime):
            raise ValueError("u and obstime must both be NumPy arrays or both torch CUDA tensors")
        t_a, t_ptr, nt, _, stream = ephemeris._times(obstime)
        if nt != n:
            raise ValueError("obstime must hold one time per unit vector")
        out_dtype = np.dtype(np.float64 if out_dtype is None else out_dtype)
        shape = (self.ncomps, n) if return_comps else (n,)
        if device_mem:
            import torch

            tdtype = torch.float64 if out_dtype == np.float64 else torch.float32
            if out is None:
                out = torch.empty(shape, dtype=tdtype, device=u_a.device)
            elif tuple(out.shape) != shape or out.dtype != tdtype or not out.is_contiguous():
                raise ValueError("out has wrong shape/dtype or is not contiguous")
            out_ptr = out.data_ptr()
        else:
            if out is None:
                out = np.empty(shape, dtype=out_dtype)
            elif out.shape != shape or out.dtype != out_dtype or out.strides[-1] != out.itemsize:
                raise ValueError("out has wrong shape/dtype or its last axis is not contiguous")
            out_ptr = out.ctypes.data
        out_row_stride = out.strides[0] // out.itemsize if (not device_mem and out.ndim == 2 and n > 1) else n
        if n == 0:
            return out
        if observer == "prepared":
            if outside_flags is None:
                raise ValueError("observer='prepared' needs outside_flags")
            flags = self._checked_flags(outside_flags)
        else:
            r_max = ephemeris.prepare(t_a, observer)
            flags = spec_outside_flags(self.spec, r_max) if outside_flags is None else self._checked_flags(outside_flags)
        args = _cabi.EvalArgs()
        args.n = n
        args.u, args.u_stride = u_ptr, u_stride
        args.outside_flags = flags.ctypes.data_as(_cabi.c_uint8_p)
        args.return_comps = int(bool(return_comps))
        args.precision = _PRECISIONS[precision]
        args.out_dtype = _cabi.OUT_F64 if out_dtype == np.float64 else _cabi.OUT_F32
        args.memory = _cabi.MEM_DEVICE if device_mem else _cabi.MEM_HOST
        args.out, args.out_stride = out_ptr, out_row_stride
        args.stream = stream
        # host arrays: prepare() -> zodi_ephemeris_stats staged the times on the device; obstime = NULL
        # integrates from that copy, so they cross the bus once
        staged = not device_mem and not os.environ.get("ZODI_TOD_EXPLICIT_OBSTIME")  # env: A/B measurement only
        args.ephemeris, args.obstime = ephemeris._handle, (None if staged else t_ptr)
        self._dispatch(args, lonlat)
        return out

    def evaluate_healpix(self, nside: int, obs, earth=None, *, pix_range=None, rot=None, nest: bool = False,
                         return_comps: bool = False, precision: str = "fp64", out=None,
                         out_dtype=None, device_out: bool = False, peer_map=None):
        """Emission for HEALPix RING pixels with the directions generated ON THE DEVICE.

        Line of sight j is the centre of pixel ``pix_range[0] + j`` (default: the whole map; RING order
        unless ``nest``),
        rotated by the optional 3x3 ``rot`` (pixel frame -> mean ecliptic).  ``obs`` / ``earth``
        are single positions (3,) [AU] (instantaneous map).  Nothing but these few numbers is
        uploaded; the result is returned as a NumPy array (host; D2H pipelined inside the
        library) or, with ``device_out=True`` / ``peer_map``, left on the GPU as a torch tensor.
        """
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        nside = int(nside)
        npix = 12 * nside * nside
        lo, hi = (0, npix) if pix_range is None else (int(pix_range[0]), int(pix_range[1]))
        if not 0 <= lo <= hi <= npix:
            raise ValueError("pix_range outside the map")
        n = hi - lo
        if peer_map is not None and peer_map.cyclic is not None:
            # block-cyclic shard of [lo, hi): this rank integrates its share; the kernel maps j -> pixel
            from .sharding import cyclic_count

            n = cyclic_count(hi - lo, peer_map.cyclic[1], peer_map.cyclic[2], peer_map.cyclic[0])
        obs_h = np.ascontiguousarray(np.asarray(obs, dtype=np.float64).reshape(3, 1))
        earth_h = obs_h if earth is None else np.ascontiguousarray(
            np.asarray(earth, dtype=np.float64).reshape(3, 1))
        out_dtype = np.dtype(np.float64 if out_dtype is None else out_dtype)
        shape = (self.ncomps, n) if return_comps else (n,)
        on_device = device_out or peer_map is not None
        out_row_stride = n
        keep = []
        if on_device:
            import torch

            dev = torch.device("cuda", self.device)
            obs_d, earth_d = torch.as_tensor(obs_h, device=dev), torch.as_tensor(earth_h, device=dev)
            keep += [obs_d, earth_d]
            obs_ptr, earth_ptr = obs_d.data_ptr(), earth_d.data_ptr()
            stream = torch.cuda.current_stream(dev).cuda_stream
            if peer_map is None:
                tdtype = torch.float64 if out_dtype == np.float64 else torch.float32
                if out is None:
                    out = torch.empty(shape, dtype=tdtype, device=dev)
                elif tuple(out.shape) != shape or out.dtype != tdtype or not out.is_contiguous():
                    raise ValueError("out has wrong shape/dtype or is not contiguous")
                out_ptr = out.data_ptr()
            else:
                if peer_map.dtype != out_dtype or peer_map.rows != (self.ncomps if return_comps else 1):
                    raise ValueError("peer_map dtype/rows do not match this call")
                out_ptr, out = None, None
        else:
            obs_ptr, earth_ptr, stream = obs_h.ctypes.data, earth_h.ctypes.data, None
            if out is None:
                out = np.empty(shape, dtype=out_dtype)
            elif out.shape != shape or out.dtype != out_dtype or out.strides[-1] != out.itemsize:
                # rows may be strided (a column block of a larger (ncomps, npix) map: one device's share)
                raise ValueError("out has wrong shape/dtype or its last axis is not contiguous")
            out_ptr = out.ctypes.data
            if out.ndim == 2 and n > 1:
                out_row_stride = out.strides[0] // out.itemsize
        if n == 0:
            return out
        if peer_map is not None:
            if peer_map.cyclic is not None and hi - lo != peer_map.n_total - peer_map.offset:
                raise ValueError("with a block-cyclic peer map pix_range must span exactly the mapped pixels")
            self._check_peer_slice(peer_map, n)
        flags = spec_outside_flags(self.spec, float(np.sqrt((obs_h ** 2).sum())))
        h = _cabi.HealpixArgs()
        a = h.base
        a.n = n
        a.u, a.u_stride = None, 0
        a.obs, a.n_obs, a.obs_stride = obs_ptr, 1, 1
        a.earth, a.n_earth, a.earth_stride = earth_ptr, 1, 1
        a.outside_flags = flags.ctypes.data_as(_cabi.c_uint8_p)
        a.return_comps = int(bool(return_comps))
        a.precision = _PRECISIONS[precision]
        a.out_dtype = _cabi.OUT_F64 if out_dtype == np.float64 else _cabi.OUT_F32
        a.memory = _cabi.MEM_DEVICE if on_device else _cabi.MEM_HOST
        a.out, a.out_stride = out_ptr, out_row_stride
        a.stream = stream
        if peer_map is not None:
            a.n_peers = len(peer_map.pointers)
            for i, ptr in enumerate(peer_map.pointers):
                a.peer_out[i] = ptr
            a.peer_offset, a.peer_stride = peer_map.offset, peer_map.n_total
            if peer_map.cyclic is not None:
                a.cyclic_block, a.cyclic_parts, a.cyclic_rank = peer_map.cyclic
        h.nside, h.ipix_start, h.nest = nside, lo, int(bool(nest))
        if rot is not None:
            r = np.asarray(rot, dtype=np.float64).reshape(9)
            h.has_rot = 1
            for i in range(9):
                h.rot[i] = float(r[i])
        _cabi.check(self._lib.zodi_evaluate_healpix(self._handle, C.byref(h)))
        del keep
        return out

    def number_density(self, xyz, earth) -> np.ndarray:
        """Number density of every component at heliocentric points ``xyz`` (3, n) [AU] for one
        Earth position (the array part of ``grid_number_density``); returns (ncomps, n) float64."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        if xyz.ndim != 2 or xyz.shape[0] != 3:
            raise ValueError("xyz must have shape (3, n)")
        earth = np.ascontiguousarray(np.asarray(earth, dtype=np.float64).reshape(3))
        n = xyz.shape[1]
        out = np.empty((self.ncomps, n), dtype=np.float64)
        _cabi.check(self._lib.zodi_number_density(self._handle, xyz.ctypes.data, n, max(n, 1),
                                                  _cabi.as_double_p(earth), out.ctypes.data, max(n, 1),
                                                  _cabi.MEM_HOST, None))
        return out

    @property
    def kernel_name(self) -> str:
        return self._lib.zodi_model_kernel_name(self._handle).decode()

    def kernel_name_for(self, n: int, precision: str = "fp64") -> str:
        """The exact kernel an evaluation of ``n`` lines of sight would launch."""
        return self._lib.zodi_model_kernel_for(self._handle, int(n), _PRECISIONS[precision]).decode()

    def last_kernel_ms(self) -> float:
        return float(self._lib.zodi_last_kernel_ms(self._handle))


class MultiDeviceModel:
    """One model replicated on several GPUs of the box, driven from ONE process for host arrays.

    The GPU counterpart of the reference's ``nprocesses`` fork pool (``zodipy/model.py:182-198``) for a
    caller that is not launched with one process per GPU: the lines of sight are split into
    ``np.array_split`` chunks (per-sample observer / Earth arrays likewise, single positions and the
    parameters replicated), every chunk is integrated by its own device through the library's pipelined
    host path (each GPU moves its chunk over its own PCIe link), and the results land in place in the
    one output array - the ``np.concatenate`` of the reference without a copy.  The early-out flags are
    formed once from ALL observers (quirk Q1), so the result is bit-identical to a single-device call.
    ctypes releases the GIL during the calls, so plain Python threads keep all devices busy.
    """

    MIN_LOS_PER_DEVICE = 1 << 15  # below this per device the split costs more than it saves

    def __init__(self, spec: dict, devices):
        devices = [int(d) for d in devices]
        if not devices:
            raise ValueError("devices must not be empty")
        self.devices = devices
        self.models = [DeviceModel(spec, d) for d in devices]
        self.spec = spec
        self.ncomps = self.models[0].ncomps
        self._pool = None  # worker threads, one per device, kept across calls

    def _workers(self):
        from concurrent.futures import ThreadPoolExecutor

        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=len(self.models), thread_name_prefix="zodi-dev")
        return self._pool

    def update(self, spec: dict) -> None:
        for m in self.models:
            m.update(spec)
        self.spec = spec
        self.ncomps = self.models[0].ncomps

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for m in self.models:
            m.close()

    def _run(self, n, obs, earth, return_comps, out, out_dtype, outside_flags, call):
        from .sharding import split_bounds

        out_dtype = np.dtype(np.float64 if out_dtype is None else out_dtype)
        shape = (self.ncomps, n) if return_comps else (n,)
        if out is None:
            out = np.empty(shape, dtype=out_dtype)
        elif out.shape != shape or out.dtype != out_dtype or not out.flags.c_contiguous:
            raise ValueError("out has wrong shape/dtype or is not C-contiguous")
        if n == 0:
            return out
        obs = np.asarray(obs, dtype=np.float64).reshape(3, -1)
        earth = obs if earth is None else np.asarray(earth, dtype=np.float64).reshape(3, -1)
        if obs.shape[1] not in (1, n) or earth.shape[1] not in (1, n):
            raise ValueError("obs/earth must hold one position or one per line of sight")
        flags = self.models[0].outside_flags(obs) if outside_flags is None else outside_flags
        # small jobs are latency-bound: one device, no thread hand-off
        parts = len(self.models) if n >= self.MIN_LOS_PER_DEVICE * len(self.models) else 1
        shards = [(m, lo, hi) for m, (lo, hi) in zip(self.models, split_bounds(n, parts)) if hi > lo]
        if len(shards) == 1:
            call(shards[0][0], 0, n, obs, earth, out, flags)
            return out

        def work(item):
            m, lo, hi = item
            part = lambda a: a[:, lo:hi] if a.shape[1] == n and n > 1 else a  # noqa: E731
            call(m, lo, hi, part(obs), part(earth), out[..., lo:hi], flags)

        list(self._workers().map(work, shards))  # re-raises the first worker exception
        return out

    def evaluate(self, u, obs, earth=None, *, return_comps: bool = False, precision: str = "fp64", out=None,
                 out_dtype=None, outside_flags=None):
        """Like :meth:`DeviceModel.evaluate` for NumPy inputs, sharded over ``devices``."""
        u = np.asarray(u, dtype=np.float64)
        if u.ndim != 2 or u.shape[0] != 3:
            raise ValueError("unit_vectors must have shape (3, n)")
        n = u.shape[1]

        def call(m, lo, hi, o, e, out_part, flags):
            m.evaluate(u[:, lo:hi], o, e, return_comps=return_comps, precision=precision, out=out_part,
                       out_dtype=out_part.dtype, outside_flags=flags)

        return self._run(n, obs, earth, return_comps, out, out_dtype, outside_flags, call)

    def evaluate_lonlat(self, lon, lat, obs, earth=None, *, rot=None, return_comps: bool = False,
                        precision: str = "fp64", out=None, out_dtype=None, outside_flags=None):
        """Like :meth:`DeviceModel.evaluate_lonlat` for NumPy inputs, sharded over ``devices``."""
        lon = np.ascontiguousarray(lon, dtype=np.float64).reshape(-1)
        lat = np.ascontiguousarray(lat, dtype=np.float64).reshape(-1)
        if lon.size != lat.size:
            raise ValueError("lon and lat must have the same length")

        def call(m, lo, hi, o, e, out_part, flags):
            m.evaluate_lonlat(lon[lo:hi], lat[lo:hi], o, e, rot=rot, return_comps=return_comps, precision=precision,
                              out=out_part, out_dtype=out_part.dtype, outside_flags=flags)

        return self._run(lon.size, obs, earth, return_comps, out, out_dtype, outside_flags, call)


    def evaluate_healpix(self, nside: int, obs, earth=None, *, pix_range=None, rot=None, nest: bool = False,
                         return_comps: bool = False, precision: str = "fp64", out=None, out_dtype=None):
        """Like :meth:`DeviceModel.evaluate_healpix` (host output), the pixel range split over ``devices``
        by the ``np.array_split`` rule: every GPU generates its own directions, integrates them and writes
        its share of the map straight into ``out`` over its own PCIe link (only the map crosses the bus)."""
        nside = int(nside)
        npix = 12 * nside * nside
        lo, hi = (0, npix) if pix_range is None else (int(pix_range[0]), int(pix_range[1]))
        if not 0 <= lo <= hi <= npix:
            raise ValueError("pix_range outside the map")

        def call(m, a, b, o, e, out_part, flags):
            m.evaluate_healpix(nside, o, e, pix_range=(lo + a, lo + b), rot=rot, nest=nest, return_comps=return_comps,
                               precision=precision, out=out_part, out_dtype=out_part.dtype)

        obs = np.asarray(obs, dtype=np.float64).reshape(3, -1)
        if obs.shape[1] != 1:
            raise ValueError("evaluate_healpix takes a single observer position")
        return self._run(hi - lo, obs, earth, return_comps, out, out_dtype, np.zeros((self.ncomps, 2), np.uint8), call)

    def ephemeris(self, t0: float, dt: float, earth_knots, obs_knots=None) -> "MultiDeviceEphemeris":
        """The same ephemeris splines on every device of this model (for :meth:`evaluate_tod`)."""
        return MultiDeviceEphemeris(t0, dt, earth_knots, obs_knots, self.devices)

    def evaluate_tod(self, u, obstime, ephemeris: "MultiDeviceEphemeris", *, observer: str = "earth", lonlat=None,
                     rot=None, return_comps: bool = False, precision: str = "fp64", out=None, out_dtype=None):
        """Time-ordered data with on-device ephemerides, the samples split over ``devices``.

        ``u`` (3, N) unit vectors, or ``lonlat=(lon, lat)`` [rad] with the optional frame rotation ``rot``;
        ``obstime`` (N,).  The reductions the reference's semantics need over ALL samples (sum |earth|^2 for
        the semb-l2 norm, quirk Q5; largest observer distance for the early-out flags, quirk Q1) are formed
        per device by ``zodi_ephemeris_stats`` - which also stages each shard's times on its device - and
        combined on the host before any device integrates, so the result is bit-identical to one device.
        """
        from .sharding import split_bounds

        if ephemeris.devices != self.devices:
            raise ValueError("ephemeris was built for other devices")
        t = np.ascontiguousarray(obstime, dtype=np.float64).reshape(-1)
        n = t.size
        if lonlat is None:
            u = np.asarray(u, dtype=np.float64)
            if u.ndim != 2 or u.shape != (3, n):
                raise ValueError("unit_vectors must have shape (3, n) with one obstime per vector")
        else:
            lon = np.ascontiguousarray(lonlat[0], dtype=np.float64).reshape(-1)
            lat = np.ascontiguousarray(lonlat[1], dtype=np.float64).reshape(-1)
            if lon.size != n or lat.size != n:
                raise ValueError("lon / lat / obstime must have the same length")
        out_dtype = np.dtype(np.float64 if out_dtype is None else out_dtype)
        shape = (self.ncomps, n) if return_comps else (n,)
        if out is None:
            out = np.empty(shape, dtype=out_dtype)
        elif out.shape != shape or out.dtype != out_dtype or not out.flags.c_contiguous:
            raise ValueError("out has wrong shape/dtype or is not C-contiguous")
        if n == 0:
            return out
        parts = len(self.models) if n >= self.MIN_LOS_PER_DEVICE * len(self.models) else 1
        shards = [(k, lo, hi) for k, (lo, hi) in enumerate(split_bounds(n, parts)) if hi > lo]
        if observer == "knots" and not ephemeris.has_obs_knots:
            raise ValueError("this ephemeris has no observer knots")
        if observer in ("earth", "semb-l2") and ephemeris.has_obs_knots:
            raise ValueError("this ephemeris carries observer knots; use observer='knots'")
        if observer not in ("earth", "semb-l2", "knots"):
            raise ValueError("observer must be 'earth', 'semb-l2' or 'knots'")
        pool = self._workers()
        for k, _, _ in shards:
            ephemeris.parts[k].set_obs_scale(1.0)
        stats = list(pool.map(lambda sh: ephemeris.parts[sh[0]].stats(t[sh[1]:sh[2]]), shards))
        sum_r2 = float(np.sum([s[0] for s in stats]))
        scale = 1.0
        if observer == "semb-l2":
            norm = float(np.sqrt(sum_r2))
            scale = (norm + MEAN_DIST_TO_L2) / norm
        r_max = max(s[2] for s in stats) if observer == "knots" else scale * max(s[1] for s in stats)
        flags = spec_outside_flags(self.spec, r_max)

        def work(sh):
            k, lo, hi = sh
            eph, m = ephemeris.parts[k], self.models[k]
            eph.set_obs_scale(scale)
            ll = None if lonlat is None else _LonLat(lon[lo:hi], lat[lo:hi], rot)
            m.evaluate(None if ll is not None else u[:, lo:hi], ephemeris=eph, obstime=t[lo:hi], observer="prepared",
                       lonlat=ll, return_comps=return_comps, precision=precision, out=out[..., lo:hi],
                       out_dtype=out_dtype, outside_flags=flags)

        list(pool.map(work, shards))
        return out


class MultiDeviceEphemeris:
    """One :class:`DeviceEphemeris` per device of a :class:`MultiDeviceModel` (same knots everywhere)."""

    def __init__(self, t0: float, dt: float, earth_knots, obs_knots=None, devices=(0,)):
        self.devices = [int(d) for d in devices]
        self.parts = [DeviceEphemeris(t0, dt, earth_knots, obs_knots, device=d) for d in self.devices]
        self.has_obs_knots = obs_knots is not None

    def close(self) -> None:
        for p in self.parts:
            p.close()


class DeviceMultiBand(DeviceModel):
    """Several bands of ONE Kelsall-family model evaluated in a single pass (shared geometry,
    temperature, table index and densities; per band one table read and a weighted sum).

    ``specs``: one neutral spec per band (``zodipy_b200.spec.build_spec`` of the same model at each
    wavelength / bandpass).  The evaluate calls of :class:`DeviceModel` apply unchanged and return
    the component-summed emission per band, shape (n_bands, N); ``return_comps`` is not available.
    """

    def __init__(self, specs, device: int = 0):
        self._lib = _cabi.load()
        if not 1 <= len(specs) <= _cabi.MAX_BANDS:
            raise ValueError(f"number of bands must be in [1, {_cabi.MAX_BANDS}]")
        self.spec = specs[0]
        self.specs = list(specs)
        self.device = int(device)
        self.n_bands = len(specs)
        self.ncomps = self.n_bands  # rows of the output
        self.n_model_comps = len(specs[0]["comps"])  # rows of outside_flags
        descs = (_cabi.ModelDesc * self.n_bands)()
        keep = []
        for i, sp in enumerate(specs):
            d, k = pack_desc(sp)
            descs[i] = d
            keep.append(k)
        self._mb = C.c_void_p()
        _cabi.check(self._lib.zodi_multiband_create(descs, self.n_bands, self.device, C.byref(self._mb)))
        del keep
        # the multi-band handle wraps a model handle as its first member
        self._handle = C.c_void_p(C.cast(self._mb, C.POINTER(C.c_void_p))[0])

    def update(self, spec):
        raise NotImplementedError("rebuild the DeviceMultiBand after a parameter update")

    def close(self) -> None:
        if getattr(self, "_mb", None) is not None and self._mb.value:
            self._lib.zodi_multiband_destroy(self._mb)
            self._mb = C.c_void_p()
            self._handle = C.c_void_p()

    def evaluate(self, u, obs=None, earth=None, **kwargs):
        kwargs.pop("return_comps", None)
        return super().evaluate(u, obs, earth, return_comps=True, **kwargs)

    def evaluate_healpix(self, nside, obs, earth=None, **kwargs):
        kwargs.pop("return_comps", None)
        return super().evaluate_healpix(nside, obs, earth, return_comps=True, **kwargs)

    def evaluate_lonlat(self, lon, lat, obs=None, earth=None, *, rot=None, **kwargs):
        kwargs.pop("return_comps", None)
        return DeviceModel.evaluate(self, None, obs, earth, lonlat=_LonLat(lon, lat, rot), return_comps=True,
                                    **kwargs)


MATH_OPS = {"log2_f64": 0, "exp2_f64": 1, "rsqrt_f64": 2, "atan2_abs_f64": 3, "asin_f32": 4,
            "atan2_abs_f32": 5, "one_minus_exp2_neg_f32": 6, "exp2_f32": 7, "log2_f32": 8}


def device_math(op: str, x, aux: float = 0.0, device: int = 0) -> np.ndarray:
    """Element-wise device math routine ``op`` (see MATH_OPS) evaluated ON the GPU (test hook)."""
    lib = _cabi.load()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    y = np.empty_like(x)
    _cabi.check(lib.zodi_device_math(int(device), MATH_OPS[op], x.size, _cabi.as_double_p(x), float(aux),
                                     _cabi.as_double_p(y)))
    return y


def kernel_launch_count() -> int:
    return int(_cabi.load().zodi_kernel_launch_count())


def healpix_vectors(nside: int, pix_range=None, rot=None, device: int = 0, nest: bool = False) -> np.ndarray:
    """(3, n) pixel-centre unit vectors (RING or NESTED) computed by the device routine (host array)."""
    nside = int(nside)
    lo, hi = (0, 12 * nside * nside) if pix_range is None else (int(pix_range[0]), int(pix_range[1]))
    out = np.empty((3, hi - lo), dtype=np.float64)
    rot_p = None
    if rot is not None:
        rot_a = np.ascontiguousarray(np.asarray(rot, dtype=np.float64).reshape(9))
        rot_p = _cabi.as_double_p(rot_a)
    _cabi.check(_cabi.load().zodi_healpix_vectors(int(device), nside, int(bool(nest)), lo, hi - lo, rot_p, out.ctypes.data,
                                                  max(hi - lo, 1), _cabi.MEM_HOST, None))
    return out


def lonlat_vectors(lon, lat, rot=None, device: int = 0) -> np.ndarray:
    """(3, n) unit vectors the lon / lat entry integrates along, computed by the device routine."""
    ll = _LonLat(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64), rot)
    out = np.empty((3, ll.n), dtype=np.float64)
    rot_p = None if ll.rot is None else _cabi.as_double_p(ll.rot)
    _cabi.check(_cabi.load().zodi_lonlat_vectors(int(device), ll.lon_ptr, ll.lat_ptr, ll.n, rot_p, out.ctypes.data,
                                                 max(ll.n, 1), _cabi.MEM_HOST, None))
    return out


def peak_probe(kind: str, device: int = 0) -> float:
    """Measured pipe peak on ``device``: 'fp32' / 'fp64' [flop/s], 'mufu' [op/s], 'hbm' [B/s]."""
    kinds = {"fp32": _cabi.PEAK_FP32_FMA, "fp64": _cabi.PEAK_FP64_FMA,
             "mufu": _cabi.PEAK_MUFU_EX2, "hbm": _cabi.PEAK_HBM_COPY}
    out = C.c_double(0.0)
    _cabi.check(_cabi.load().zodi_peak_probe(int(device), kinds[kind], C.byref(out)))
    return out.value
