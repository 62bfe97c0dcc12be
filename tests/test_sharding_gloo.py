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
