"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): split rule, global early-out flags,
all-gather assembly with ragged shards.  The compute is stubbed by the oracle (tests may use it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import zodi_oracle as oracle
from helpers import golden_case
from zodipy_b200 import sharding
from zodipy_b200.spec import outside_flags as spec_outside_flags


def test_split_bounds_is_array_split_rule():
    for n in (0, 1, 7, 10, 600, 12 * 64 * 64):
        for parts in (1, 2, 3, 4, 8):
            ref = np.array_split(np.arange(n), parts)
            got = sharding.split_bounds(n, parts)
            assert len(got) == parts
            for (lo, hi), r in zip(got, ref):
                assert hi - lo == r.size and (r.size == 0 or (r[0] == lo and r[-1] == hi - 1))
            assert sharding.padded_count(n, parts) == max(hi - lo for lo, hi in got)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_with_flags(spec):
    """Oracle evaluation that takes the early-out flags as an argument instead of deriving them
    from the local observers (what a shard must do)."""

    def evaluate(u, obs, earth, outside_flags, return_comps=True):
        u, obs, earth = (np.asarray(a) for a in (u, obs, earth))
        n = u.shape[1]
        out = np.zeros((len(spec["comps"]), n))
        step = oracle.kelsall_step if spec["kind"] == "kelsall" else oracle.rrm_step
        for ci, comp in enumerate(spec["comps"]):
            # flagged bound -> eps for every line of sight, like sphere_intersection's early-out
            rng = []
            for b in (0, 1):
                if outside_flags[ci, b]:
                    rng.append(np.full(max(obs.shape[1], 1), oracle.EPS))
                else:
                    rng.append(oracle.sphere_intersection(obs, u, comp["cutoff"][b]))
            acc = 0
            for x, w in zip(spec["points"], spec["weights"]):
                acc = acc + step(x, rng[0], rng[1], obs, u, earth, spec, comp) * w
            out[ci] = acc
        return torch.from_numpy(out if return_comps else out.sum(axis=0))

    return evaluate


def _worker(rank, world, port, case_id, per_sample, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, a = golden_case(case_id)
        spec = case["spec"]
        n = a["u"].shape[1]
        lo, hi = sharding.split_bounds(n, world)[rank]
        obs = a["obs"][:, lo:hi] if per_sample else a["obs"]
        earth = a["earth"][:, lo:hi] if per_sample else a["earth"]
        out = sharding.evaluate_sharded(
            _oracle_with_flags(spec),
            lambda o: float(np.sqrt((np.asarray(o) ** 2).sum(axis=0)).max()),
            lambda r: spec_outside_flags(spec, r),
            a["u"][:, lo:hi], obs, earth, n, obs_per_sample=per_sample, return_comps=True)
        queue.put((rank, out.numpy()))
    except Exception as err:  # surface worker failures immediately instead of a queue timeout
        queue.put((rank, err))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case_id,per_sample", [("dirbe_25um_tod_straddle", True),
                                                ("planck18_857", False)])
def test_sharded_equals_single_rank_bitwise(world, case_id, per_sample):
    """Sharded result (any world size, ragged shards) is bit-identical to the one-process result,
    including the case where only SOME shards contain observers beyond a cutoff (quirk Q1)."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_id, per_sample, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=180) for _ in range(world))
    for r, v in results.items():
        assert not isinstance(v, Exception), f"rank {r}: {v!r}"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, a = golden_case(case_id)
    for r in range(world):
        np.testing.assert_array_equal(results[r], a["emission"])  # every rank holds the full map


@pytest.mark.parametrize("n_devices", [2, 3])
@pytest.mark.parametrize("case_id,per_sample", [("dirbe_25um_tod_straddle", True), ("planck18_857", False)])
def test_single_process_multi_device_host_sharding(monkeypatch, n_devices, case_id, per_sample):
    """MultiDeviceModel (one process driving several GPUs for host arrays): array_split shards,
    flags formed once from ALL observers, results written in place into one output - bit-identical to
    the unsharded evaluation.  The per-device compute is stubbed by the oracle."""
    from zodipy_b200 import engine

    case, a = golden_case(case_id)
    spec = case["spec"]
    calls = []

    class StubDeviceModel:
        def __init__(self, spec_, device):
            self.spec, self.device, self.ncomps = spec_, device, len(spec_["comps"])
            self._eval = _oracle_with_flags(spec_)

        def outside_flags(self, obs):
            return spec_outside_flags(self.spec, float(np.sqrt((np.asarray(obs) ** 2).sum(axis=0)).max()))

        def evaluate(self, u, obs, earth, *, return_comps, precision, out, out_dtype, outside_flags):
            calls.append((self.device, u.shape[1], obs.shape[1]))
            assert out.strides[-1] == out.itemsize and out.dtype == out_dtype
            out[...] = self._eval(u, obs, earth, outside_flags, return_comps).numpy()

        def evaluate_lonlat(self, lon, lat, obs, earth, *, rot, **kw):
            u = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
            self.evaluate(u if rot is None else np.asarray(rot).reshape(3, 3) @ u, obs, earth, **kw)

        def update(self, spec_):
            self.spec = spec_

        def close(self):
            pass

    monkeypatch.setattr(engine, "DeviceModel", StubDeviceModel)
    monkeypatch.setattr(engine.MultiDeviceModel, "MIN_LOS_PER_DEVICE", 1)  # the golden cases are small
    multi = engine.MultiDeviceModel(spec, list(range(n_devices)))
    n = a["u"].shape[1]
    got = multi.evaluate(a["u"], a["obs"], a["earth"], return_comps=True)
    np.testing.assert_array_equal(got, a["emission"])
    assert sorted(c[1] for c in calls) == sorted(hi - lo for lo, hi in sharding.split_bounds(n, n_devices))
    assert all(c[2] == (c[1] if per_sample else 1) for c in calls)
    out = np.full(n, np.nan)
    assert multi.evaluate(a["u"], a["obs"], a["earth"], out=out) is out
    np.testing.assert_array_equal(out, a["emission"].sum(axis=0))
    # spherical-coordinate entry: same shards, angles split instead of vectors
    lon, lat = np.arctan2(a["u"][1], a["u"][0]), np.arcsin(np.clip(a["u"][2], -1, 1))
    got_ll = multi.evaluate_lonlat(lon, lat, a["obs"], a["earth"], return_comps=True)
    np.testing.assert_allclose(got_ll, a["emission"], rtol=1e-9, atol=1e-30)
    # fewer lines of sight than devices: empty shards are skipped
    few = multi.evaluate(a["u"][:, :1], a["obs"][:, :1], a["earth"][:, :1])
    assert few.shape == (1,) and np.isfinite(few).all()
    assert multi.evaluate(np.empty((3, 0)), a["obs"][:, :1]).shape == (0,)
    with pytest.raises(ValueError):
        multi.evaluate(a["u"], a["obs"][:, :2] if per_sample else np.ones((3, 2)))
    with pytest.raises(ValueError):
        engine.MultiDeviceModel(spec, [])
    # small jobs stay on the first device
    monkeypatch.setattr(engine.MultiDeviceModel, "MIN_LOS_PER_DEVICE", 1 << 15)
    calls.clear()
    np.testing.assert_array_equal(multi.evaluate(a["u"], a["obs"], a["earth"], return_comps=True), a["emission"])
    assert [c[0] for c in calls] == [0]


@pytest.mark.parametrize("observer", ["earth", "semb-l2"])
def test_single_process_multi_device_tod_and_healpix(monkeypatch, observer):
    """MultiDeviceModel.evaluate_tod: the reductions over ALL samples (semb-l2 whole-array norm, quirk Q5;
    global early-out flags, quirk Q1) are combined on the host before any device integrates, every shard
    then uses the global observer scale - equal to the unsharded evaluation.  evaluate_healpix: pixel
    ranges split by the array_split rule, results in place.  Per-device compute is stubbed by the oracle
    (host CubicSpline positions)."""
    from scipy.interpolate import CubicSpline

    from zodipy_b200 import engine, healpix

    case, _ = golden_case("dirbe_25um_rand")
    spec = case["spec"]
    t0, dt, n_knots = 59215.0, 1.0 / 24.0, 60 * 24
    tk = t0 + dt * np.arange(n_knots)
    lon_e = 2 * np.pi * (tk - t0) / 365.25 + 1.7
    earth_knots = (1.0 - 0.0167 * np.cos(lon_e - 1.8)) * np.array([np.cos(lon_e), np.sin(lon_e), 1e-5 * np.sin(3 * lon_e)])
    spline = CubicSpline(tk, earth_knots, axis=-1)

    class StubEphemeris:
        def __init__(self, t0_, dt_, earth_knots_, obs_knots_=None, device=0):
            self.device, self.scale, self.staged = device, 1.0, None

        def set_obs_scale(self, s):
            self.scale = float(s)

        def stats(self, t):
            e = spline(np.asarray(t))
            r2 = (e * e).sum(axis=0)
            self.staged = np.array(t)
            return float(r2.sum()), float(np.sqrt(r2.max())), self.scale * float(np.sqrt(r2.max()))

        def close(self):
            pass

    class StubDeviceModel:
        def __init__(self, spec_, device):
            self.spec, self.device, self.ncomps = spec_, device, len(spec_["comps"])
            self._eval = _oracle_with_flags(spec_)

        def evaluate(self, u, obs=None, earth=None, *, ephemeris=None, obstime=None, observer="earth", lonlat=None,
                     return_comps=False, precision="fp64", out=None, out_dtype=None, outside_flags=None):
            assert observer == "prepared" and np.array_equal(ephemeris.staged, obstime)  # stats ran on THIS shard
            if lonlat is not None:
                u = np.array([np.cos(lonlat.lat) * np.cos(lonlat.lon), np.cos(lonlat.lat) * np.sin(lonlat.lon),
                              np.sin(lonlat.lat)])
            e = spline(obstime)
            out[...] = self._eval(u, ephemeris.scale * e, e, outside_flags, return_comps).numpy()

        def evaluate_healpix(self, nside, obs, earth=None, *, pix_range, rot, nest, return_comps, precision, out,
                             out_dtype):
            u = healpix.pix2vec_ring(nside, np.arange(*pix_range))
            flags = spec_outside_flags(self.spec, float(np.sqrt((np.asarray(obs) ** 2).sum())))
            out[...] = self._eval(u, np.asarray(obs).reshape(3, 1), np.asarray(obs if earth is None else earth).reshape(3, 1),
                                  flags, return_comps).numpy()

        def close(self):
            pass

    monkeypatch.setattr(engine, "DeviceModel", StubDeviceModel)
    monkeypatch.setattr(engine, "DeviceEphemeris", StubEphemeris)
    monkeypatch.setattr(engine.MultiDeviceModel, "MIN_LOS_PER_DEVICE", 1)
    multi = engine.MultiDeviceModel(spec, [0, 1, 2])
    rng = np.random.default_rng(3)
    n = 301
    t = np.sort(rng.uniform(tk[0], tk[-1], n))
    u = rng.normal(size=(3, n))
    u /= np.linalg.norm(u, axis=0)
    eph = multi.ephemeris(t0, dt, earth_knots)
    got = multi.evaluate_tod(u, t, eph, observer=observer, return_comps=True)
    earth = spline(t)
    norm = np.linalg.norm(earth)  # un-axised: the whole (3, n) array (zodipy/bodies.py:47)
    scale = (norm + engine.MEAN_DIST_TO_L2) / norm if observer == "semb-l2" else 1.0
    ref = oracle.evaluate(spec, u, scale * earth, earth)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-300)
    lon, lat = np.arctan2(u[1], u[0]), np.arcsin(np.clip(u[2], -1, 1))
    got_ll = multi.evaluate_tod(None, t, eph, observer=observer, lonlat=(lon, lat))
    np.testing.assert_allclose(got_ll, ref.sum(axis=0), rtol=1e-9)
    with pytest.raises(ValueError):
        multi.evaluate_tod(u, t[:-1], eph)
    # HEALPix map split over the devices
    nside = 4
    obs = np.array([-0.39, 0.90, 0.0])
    hp = multi.evaluate_healpix(nside, obs, return_comps=True)
    ref_hp = oracle.evaluate(spec, healpix.pix2vec_ring(nside, np.arange(12 * nside * nside)), obs.reshape(3, 1),
                             obs.reshape(3, 1))
    np.testing.assert_array_equal(hp, ref_hp)
    part = multi.evaluate_healpix(nside, obs, pix_range=(10, 77))
    np.testing.assert_array_equal(part, ref_hp.sum(axis=0)[10:77])


def test_peer_slice_bounds_are_checked_before_launch():
    """The kernel's peer stores go to remote GPU memory: a slice that does not fit the mapped maps must be
    refused in Python (contiguous and block-cyclic layouts), whatever return_comps is."""
    from types import SimpleNamespace

    from zodipy_b200 import engine

    check = engine.DeviceModel._check_peer_slice
    contiguous = SimpleNamespace(cyclic=None, offset=100, n_total=1000)
    check(None, contiguous, 900)
    with pytest.raises(ValueError):
        check(None, contiguous, 901)
    n_total, parts, block = 1000, 3, 64
    for rank in range(parts):
        pm = SimpleNamespace(cyclic=(block, parts, rank), offset=0, n_total=n_total)
        mine = sharding.cyclic_count(n_total, parts, rank, block)
        check(None, pm, mine)
        for wrong in (mine - 1, mine + 1, n_total):
            if wrong != mine:
                with pytest.raises(ValueError):
                    check(None, pm, wrong)
    assert sum(sharding.cyclic_count(n_total, parts, r, block) for r in range(parts)) == n_total
