"""``Model``: the reference's public interface over the B200 line-of-sight integrator.

Signature, argument meaning, output shapes and error behaviour follow
``zodipy/model.py:33-333``.  The Astropy-side host work of ``Model.evaluate`` (ephemerides, frame
rotation) is unchanged in spirit and needs Astropy, which is imported lazily; everything below
the array seam (``zodipy/model.py:253-279``) runs in the CUDA library.  ``evaluate_xyz`` exposes
that seam directly (no Astropy): it is what the benchmarks, the parity tests and multi-GPU
sharding use.

Additive, defaulted knobs (not in the reference): ``precision`` ("fp64" faithful | "fp32" fast),
``device`` (CUDA ordinal; default ``LOCAL_RANK`` or 0), ``tod_ephemeris`` ("host": per-sample
positions interpolated with SciPy as in the reference | "device": hourly knots uploaded once and
interpolated in the kernel prologue), ``sky_rotation`` ("device": SkyCoord longitudes /
latitudes are uploaded as they are and the rotation to the mean ecliptic happens in the kernel
prologue whenever Astropy's transformation of the frame is a fixed rotation | "host": every
coordinate is transformed by Astropy as in the reference), ``devices`` (several CUDA ordinals driven
by this one process for host arrays).  ``nprocesses=k`` of ``evaluate`` - the reference's number of
fork-pool workers (``zodipy/model.py:182-198``) - maps to ``min(k, visible GPUs)`` devices when neither
``device`` nor ``devices`` was given and the process is not one rank of a one-process-per-GPU job
(``LOCAL_RANK`` unset); the split rule (``np.array_split``) and the results are the reference's.
"""
from __future__ import annotations

import os

import numpy as np

from . import spectral
from . import units as zu
from .component import ComponentLabel
from .engine import DeviceModel
from .spec import build_spec
from .zodiacal_light_model import clone_model, model_registry


class Model:
    """Main interface (drop-in for ``zodipy.Model``)."""

    def __init__(self, x, *, weights=None, name: str = "dirbe", gauss_quad_degree: int = 50,
                 extrapolate: bool = False, ephemeris: str = "builtin",
                 precision: str = "fp64", device: int | None = None,
                 tod_ephemeris: str = "host", sky_rotation: str = "device", devices=None) -> None:
        try:
            if not x.isscalar and weights is None:
                raise ValueError("Bandpass weights must be provided for non-scalar `x`.")
        except AttributeError as error:
            raise TypeError("The input 'x' must be an astropy Quantity.") from error
        if not zu.is_quantity(x):
            raise TypeError("The input 'x' must be an astropy Quantity.")
        if x.isscalar and weights is not None:
            raise ValueError("Bandpass weights should not be provided for scalar `x`.")

        self._ipd_model = clone_model(model_registry.get_model(name))
        if not extrapolate and not self._ipd_model.is_valid_at(x):
            raise ValueError(
                "The requested frequencies are outside the valid range of the model. "
                "If this was intended, set the extrapolate argument to True.")

        if weights is not None:
            weights = np.asarray(weights, dtype=np.float64)
            if x.size != weights.size:
                raise ValueError("Number of wavelengths and weights must be the same in the bandpass.")
            # normalised over the user's x values in the user's unit (model.py:94; quirk Q6)
            normalized_weights = weights / spectral.trapezoid(weights, zu.native_value(x))
        else:
            normalized_weights = None

        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        if tod_ephemeris not in ("host", "device"):
            raise ValueError("tod_ephemeris must be 'host' or 'device'")
        if sky_rotation not in ("host", "device"):
            raise ValueError("sky_rotation must be 'host' or 'device'")
        self._tod_ephemeris = tod_ephemeris
        self._sky_rotation = sky_rotation
        self._x = x
        self._bounds_error = not extrapolate
        self._normalized_weights = normalized_weights
        self._gauss_quad_degree = int(gauss_quad_degree)
        self._ephemeris = ephemeris
        self._precision = precision
        if devices is not None:
            devices = [int(d) for d in devices]
            if not devices or (device is not None and int(device) != devices[0]):
                raise ValueError("devices must be a non-empty list whose first entry equals device (if given)")
            device = devices[0]
        # nprocesses -> devices only when the caller left the placement to us (see module docstring)
        self._auto_devices = device is None and devices is None and "LOCAL_RANK" not in os.environ
        self._device = int(os.environ.get("LOCAL_RANK", 0)) if device is None else int(device)
        self._devices = devices if devices is not None else [self._device]
        self._device_model: DeviceModel | None = None
        self._multi_model = None
        self._init_ipd_model_partials()

    # ---------------------------------------------------------------------------------------
    def _init_ipd_model_partials(self) -> None:
        """(Re)build the device parameter block source (``zodipy/model.py:281-301``)."""
        self._spec = build_spec(self._ipd_model, self._x, self._normalized_weights,
                                self._bounds_error, self._gauss_quad_degree)
        self._b_nu_table = self._spec["table"]
        if self._device_model is not None:
            self._device_model.update(self._spec)
        if self._multi_model is not None:
            self._multi_model.update(self._spec)

    @property
    def spec(self) -> dict:
        """The neutral model specification uploaded to the device (see ``zodipy_b200.spec``)."""
        return self._spec

    @property
    def ncomps(self) -> int:
        return self._ipd_model.ncomps

    @property
    def device_model(self) -> DeviceModel:
        if self._device_model is None:
            self._device_model = DeviceModel(self._spec, self._device)
        return self._device_model

    def _multi(self, *arrays):
        """The multi-GPU engine when ``devices`` names several GPUs and the inputs are host arrays."""
        if len(self._devices) < 2 or any(type(a).__module__.startswith("torch") for a in arrays):
            return None
        if self._multi_model is None:
            from .engine import MultiDeviceModel

            self._multi_model = MultiDeviceModel(self._spec, self._devices)
        return self._multi_model

    def _use_processes(self, nprocesses: int) -> None:
        """``nprocesses=k`` -> ``min(k, visible GPUs)`` devices (``zodipy/model.py:182-198`` splits the
        coordinates over k workers; here the workers are GPUs driven by threads of this process)."""
        if not self._auto_devices or nprocesses is None or int(nprocesses) <= 1:
            return
        from . import _cabi

        try:
            visible = _cabi.device_count()
        except (_cabi.ZodiError, RuntimeError):
            return  # no usable device: the evaluation itself reports that (there is no CPU fallback)
        k = max(1, min(int(nprocesses), visible))
        wanted = list(range(k))
        if wanted != self._devices:
            if self._multi_model is not None:
                self._multi_model.close()
                self._multi_model = None
            self._devices = wanted
            self._device = wanted[0]

    def ephemeris(self, t0: float, dt: float, earth_knots, obs_knots=None):
        """Device-resident ephemeris splines for :meth:`evaluate_tod_xyz` / :meth:`evaluate_lonlat`, on
        every device this model drives (hourly knots ``t0 + k dt`` as in ``zodipy/bodies.py:16-35``)."""
        from .engine import DeviceEphemeris, MultiDeviceEphemeris

        if len(self._devices) > 1:
            return MultiDeviceEphemeris(t0, dt, earth_knots, obs_knots, self._devices)
        return DeviceEphemeris(t0, dt, earth_knots, obs_knots, device=self._device)

    # ---------------------------------------------------------------------------------------
    def evaluate_xyz(self, unit_vectors, obs_xyz, earth_xyz=None, *, return_comps: bool = False,
                     precision: str | None = None, out=None, out_dtype=None, outside_flags=None):
        """The array seam (``zodipy/model.py:253-279,202-203``) without Astropy.

        unit_vectors: (3, N) ecliptic unit vectors; obs_xyz / earth_xyz: (3,), (3, 1) or (3, N)
        heliocentric ecliptic positions [AU] (``earth_xyz`` defaults to ``obs_xyz``).  NumPy in ->
        NumPy out; torch CUDA tensors in -> torch CUDA tensor out.  Values are MJy/sr.
        """
        multi = self._multi(unit_vectors)
        if multi is not None:
            return multi.evaluate(unit_vectors, obs_xyz, earth_xyz, return_comps=return_comps,
                                  precision=precision or self._precision, out=out, out_dtype=out_dtype,
                                  outside_flags=outside_flags)
        return self.device_model.evaluate(
            unit_vectors, obs_xyz, earth_xyz, return_comps=return_comps,
            precision=precision or self._precision, out=out, out_dtype=out_dtype,
            outside_flags=outside_flags)

    def evaluate_tod_xyz(self, unit_vectors, obstime, ephemeris, *, observer: str = "earth",
                         return_comps: bool = False, precision: str | None = None, out=None,
                         out_dtype=None):
        """Time-ordered data with Earth / observer positions interpolated ON THE DEVICE
        (additive entry).

        ``ephemeris``: a :class:`zodipy_b200.engine.DeviceEphemeris` built from positions at
        uniformly spaced knots (the reference's hourly grid, ``zodipy/bodies.py:16-35``);
        ``obstime`` (N,): per-sample times in the knots' unit; ``observer``: "earth", "semb-l2" or
        "knots".  Equivalent to ``evaluate_xyz(u, obs_xyz(t), earth_xyz(t))`` with the positions
        from ``scipy.interpolate.CubicSpline`` - without computing or uploading them on the host.
        """
        if hasattr(ephemeris, "parts"):  # MultiDeviceEphemeris: samples split over the devices
            return self._multi(unit_vectors).evaluate_tod(
                unit_vectors, obstime, ephemeris, observer=observer, return_comps=return_comps,
                precision=precision or self._precision, out=out, out_dtype=out_dtype)
        return self.device_model.evaluate(
            unit_vectors, return_comps=return_comps, precision=precision or self._precision, out=out,
            out_dtype=out_dtype, ephemeris=ephemeris, obstime=obstime, observer=observer)

    def evaluate_lonlat(self, lon, lat, obs_xyz=None, earth_xyz=None, *, frame_rotation=None,
                        return_comps: bool = False, precision: str | None = None, out=None, out_dtype=None,
                        outside_flags=None, ephemeris=None, obstime=None, observer: str = "earth"):
        """Like ``evaluate_xyz`` / ``evaluate_tod_xyz`` with the directions given as longitude /
        latitude [rad] (N,) of a frame whose constant rotation to the mean ecliptic is the 3x3
        ``frame_rotation`` (``None``: ecliptic angles).  Additive entry: the
        unit vectors of ``zodipy/model.py:247-251`` are formed and rotated in the kernel prologue,
        so the host neither builds nor uploads the (3, N) array (16 instead of 24 B per line of
        sight cross the bus)."""
        if ephemeris is not None and hasattr(ephemeris, "parts"):
            return self._multi(lon, lat).evaluate_tod(
                None, obstime, ephemeris, observer=observer, lonlat=(lon, lat), rot=frame_rotation,
                return_comps=return_comps, precision=precision or self._precision, out=out, out_dtype=out_dtype)
        multi = self._multi(lon, lat) if ephemeris is None else None
        if multi is not None:
            return multi.evaluate_lonlat(lon, lat, obs_xyz, earth_xyz, rot=frame_rotation, return_comps=return_comps,
                                         precision=precision or self._precision, out=out, out_dtype=out_dtype,
                                         outside_flags=outside_flags)
        return self.device_model.evaluate_lonlat(
            lon, lat, obs_xyz, earth_xyz, rot=frame_rotation, return_comps=return_comps,
            precision=precision or self._precision, out=out, out_dtype=out_dtype,
            outside_flags=outside_flags, ephemeris=ephemeris, obstime=obstime, observer=observer)

    def evaluate_healpix(self, nside: int, obs_xyz, earth_xyz=None, *, frame_rotation=None,
                         pix_range=None, nest: bool = False, return_comps: bool = False,
                         precision: str | None = None, out=None, out_dtype=None, device_out: bool = False):
        """Full-sky (or ``pix_range``) HEALPix map (RING, or NESTED with ``nest=True``) for one observation time, with the pixel
        directions generated on the GPU (additive entry).

        Equivalent to ``evaluate_xyz(R @ pix2vec(nside, ipix), obs_xyz, earth_xyz)`` - what the
        reference's map examples do through healpy + SkyCoord (``docs/examples/healpy_map.py``) -
        without building or uploading the (3, N) direction array.  ``frame_rotation`` is the 3x3
        matrix taking pixel-frame vectors to mean-ecliptic ones (identity if the map is ecliptic).
        """
        multi = None if device_out else self._multi()
        if multi is not None:
            return multi.evaluate_healpix(nside, obs_xyz, earth_xyz, pix_range=pix_range, rot=frame_rotation, nest=nest,
                                          return_comps=return_comps, precision=precision or self._precision, out=out,
                                          out_dtype=out_dtype)
        return self.device_model.evaluate_healpix(
            nside, obs_xyz, earth_xyz, pix_range=pix_range, rot=frame_rotation, nest=nest,
            return_comps=return_comps, precision=precision or self._precision, out=out,
            out_dtype=out_dtype, device_out=device_out)

    def evaluate(self, skycoord, *, obspos="earth", return_comps: bool = False, nprocesses: int = 1):
        """Simulated zodiacal light [MJy/sr] for an ``astropy.coordinates.SkyCoord``.

        Same contract as ``zodipy/model.py:119-203``.  Requires Astropy at call time.
        """
        try:
            if skycoord.obstime is None:
                raise ValueError("The `obstime` attribute of the `SkyCoord` object is not set.")
        except AttributeError as error:
            raise TypeError("The input coordinates must be an astropy SkyCoord object.") from error
        try:
            if not (obspos_isstr := isinstance(obspos, str)) and (
                (obspos.ndim > 1 and skycoord.obstime.size != obspos.shape[-1])
                or (obspos.ndim == 1 and skycoord.obstime.size != 1)
            ):
                raise ValueError("The number of obstime (ncoords) and obspos (3, ncoords) does not match.")
        except AttributeError as error:
            raise TypeError("The observer position is not a string or an astropy Quantity.") from error
        if skycoord.obstime.size > skycoord.size:
            raise ValueError("The size of obstime must be either 1 or ncoords.")

        self._use_processes(nprocesses)
        from . import astro  # lazy: needs astropy

        # directions: angles + one 3x3 matrix for the device, or Astropy-transformed vectors
        sky = astro.sky_lonlat_rotation(skycoord) if self._sky_rotation == "device" else None
        interp_obstimes = None
        if skycoord.obstime.size != 1:
            interp_obstimes = astro.arrange_obstimes(skycoord.obstime[0].mjd, skycoord.obstime[-1].mjd)
            if obspos_isstr and self._tod_ephemeris == "device" and interp_obstimes.size >= 4:
                # hourly knots -> device splines; per-sample interpolation in the kernel prologue
                eph, mode, u_xyz, mjd = astro.device_ephemeris(
                    skycoord, obspos, interp_obstimes, self._ephemeris, self._devices, with_directions=sky is None)
                if sky is None:
                    emission = self.evaluate_tod_xyz(u_xyz, mjd, eph, observer=mode, return_comps=return_comps)
                else:
                    emission = self.evaluate_lonlat(sky[0], sky[1], frame_rotation=sky[2], ephemeris=eph,
                                                    obstime=mjd, observer=mode, return_comps=return_comps)
                return astro.as_mjy_per_sr(emission)
        earth_xyz, obs_xyz, u_xyz = astro.prepare_arrays(
            skycoord, obspos, obspos_isstr, interp_obstimes, self._ephemeris, with_directions=sky is None)
        if sky is None:
            emission = self.evaluate_xyz(u_xyz, obs_xyz, earth_xyz, return_comps=return_comps)
        else:
            emission = self.evaluate_lonlat(sky[0], sky[1], obs_xyz, earth_xyz, frame_rotation=sky[2],
                                            return_comps=return_comps)
        return astro.as_mjy_per_sr(emission)

    # ---------------------------------------------------------------------------------------
    def get_parameters(self) -> dict:
        """Model parameter dictionary (``zodipy/model.py:303-310``)."""
        return self._ipd_model.to_dict()

    def update_parameters(self, parameters: dict) -> None:
        """Replace the model parameters (``zodipy/model.py:312-333``) and re-upload them."""
        new = parameters.copy()
        new["comps"] = {}
        for key, value in parameters.items():
            if key == "comps":
                for comp_key, comp_value in value.items():
                    label = ComponentLabel(comp_key)
                    new["comps"][label] = type(self._ipd_model.comps[label])(**comp_value)
            elif isinstance(value, dict):
                new[key] = {ComponentLabel(k): v for k, v in value.items()}
        self._ipd_model = self._ipd_model.__class__(**new)
        self._init_ipd_model_partials()


class MultiBandModel:
    """Several wavelengths / bandpasses of ONE model evaluated in a single pass over the lines of
    sight (additive API).

    ``xs``: sequence of Quantities (scalars or bandpass arrays), ``weights``: matching sequence of
    bandpass weights or ``None`` entries.  Equivalent to ``[Model(x, weights=w, name=name, ...)
    .evaluate_xyz(...) for x, w in zip(xs, weights)]`` - the per-band loop of the reference's own
    tests (tests/test_evaluate.py:54-66) - but positions, temperatures and number densities are
    computed once and shared by all bands.  Returns the component-summed emission, shape
    (n_bands, N).  Kelsall-family models (dirbe, planck13/15/18, odegard), up to 16 bands.
    """

    def __init__(self, xs, *, weights=None, name: str = "dirbe", gauss_quad_degree: int = 50,
                 extrapolate: bool = False, precision: str = "fp64", device: int | None = None) -> None:
        xs = list(xs)
        weights = [None] * len(xs) if weights is None else list(weights)
        if len(weights) != len(xs):
            raise ValueError("weights must have one entry (array or None) per band")
        # one single-band Model per band does validation, bandpass normalisation and unpacking
        self.bands = [Model(x, weights=w, name=name, gauss_quad_degree=gauss_quad_degree,
                            extrapolate=extrapolate, precision=precision, device=device)
                      for x, w in zip(xs, weights)]
        self._precision = precision
        self._device = self.bands[0]._device
        self._device_model = None

    @property
    def n_bands(self) -> int:
        return len(self.bands)

    @property
    def specs(self) -> list:
        return [m.spec for m in self.bands]

    @property
    def device_model(self):
        from .engine import DeviceMultiBand

        if self._device_model is None:
            self._device_model = DeviceMultiBand(self.specs, self._device)
        return self._device_model

    def evaluate_xyz(self, unit_vectors, obs_xyz, earth_xyz=None, *, precision: str | None = None, out=None,
                     out_dtype=None, outside_flags=None):
        """(n_bands, N) emission [MJy/sr] for ecliptic unit vectors (3, N); see ``Model.evaluate_xyz``."""
        return self.device_model.evaluate(unit_vectors, obs_xyz, earth_xyz,
                                          precision=precision or self._precision, out=out,
                                          out_dtype=out_dtype, outside_flags=outside_flags)

    def evaluate_healpix(self, nside: int, obs_xyz, earth_xyz=None, *, frame_rotation=None, pix_range=None,
                         nest: bool = False, precision: str | None = None, out=None, out_dtype=None,
                         device_out: bool = False):
        """(n_bands, npix) HEALPix maps with on-device directions; see ``Model.evaluate_healpix``."""
        return self.device_model.evaluate_healpix(
            nside, obs_xyz, earth_xyz, pix_range=pix_range, rot=frame_rotation, nest=nest,
            precision=precision or self._precision, out=out, out_dtype=out_dtype, device_out=device_out)
