"""Sharded evaluation with real kernels.

* Any box (1 GPU): two processes share cuda:0, rendezvous over gloo; the sharded map must be
  bit-identical to the single-process map (reference: nprocesses equality,
  tests/test_evaluate.py:215-262).
* Boxes with >= 2 GPUs: one process per GPU over NCCL; the fused peer-store epilogue
  (kernel writes its slice into every rank's map over NVLink) must equal the NCCL all-gather.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, backend, case_id, per_sample, precision, use_peer_map, queue):
    import torch
    import torch.distributed as dist

    from helpers import golden_case
    from zodipy_b200 import engine, sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev_index = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, a = golden_case(case_id)
        n = a["u"].shape[1]
        lo, hi = sharding.split_bounds(n, world)[rank]
        dm = engine.DeviceModel(case["spec"], device=dev_index)
        u = torch.as_tensor(np.ascontiguousarray(a["u"][:, lo:hi]), device=dev)
        obs = torch.as_tensor(np.ascontiguousarray(a["obs"][:, lo:hi] if per_sample else a["obs"]), device=dev)
        earth = torch.as_tensor(np.ascontiguousarray(a["earth"][:, lo:hi] if per_sample else a["earth"]),
                                device=dev)
        full = sharding.evaluate_sharded(
            dm.evaluate, dm.max_observer_radius, lambda r: sharding_flags(dm, r), u, obs, earth, n,
            obs_per_sample=per_sample, return_comps=True, precision=precision)
        result = {"gathered": full.cpu().numpy()}
        if use_peer_map:
            pm = sharding.PeerMap(n, dm.ncomps, np.float64, dev_index)
            r_glob = dm.max_observer_radius(obs)
            if per_sample:
                r_glob = sharding.global_max_radius(r_glob)
            dm.evaluate(u, obs, earth, return_comps=True, precision=precision,
                        outside_flags=sharding_flags(dm, r_glob), peer_map=pm)
            result["fused"] = pm.finish().cpu().numpy()
            # repeated evaluations into ONE PeerMap with different inputs and no barrier in between: the
            # double-buffered maps + the per-evaluation rendezvous must keep every map intact while peers
            # are already storing the next one (write-after-read hazard of a single-buffered map)
            maps = []
            for scale in WAR_SCALES[:2]:
                dm.evaluate(u, obs * scale, earth, return_comps=True, precision=precision,
                            outside_flags=sharding_flags(dm, r_glob * scale), peer_map=pm)
                maps.append(pm.finish())
            result["war"] = [m.cpu().numpy() for m in maps]  # read only now: evaluation 2 is already in flight
            dm.evaluate(u, obs * WAR_SCALES[2], earth, return_comps=True, precision=precision,
                        outside_flags=sharding_flags(dm, r_glob * WAR_SCALES[2]), peer_map=pm)  # reuses buffer 1
            result["war"].append(pm.finish().cpu().numpy())
            assert not pm.timed_out()
            dist.barrier()
            pm.close()
        if use_peer_map:
            # block-cyclic shards + on-device HEALPix directions + fused peer stores
            nside, block = 16, 64
            npix = 12 * nside * nside
            pmc = sharding.PeerMap(npix, dm.ncomps, np.float64, dev_index, cyclic_block=block)
            dm.evaluate_healpix(nside, a["obs"][:, :1], a["earth"][:, :1], return_comps=True,
                                precision=precision, peer_map=pmc)
            result["cyclic_healpix"] = pmc.finish().cpu().numpy()
            # array-seam inputs in block-cyclic order
            from zodipy_b200 import healpix

            idx = sharding.cyclic_indices(npix, world, rank, block)
            u_c = torch.as_tensor(healpix.pix2vec_ring(nside, idx), device=dev)
            dm.evaluate(u_c, torch.as_tensor(a["obs"][:, :1].copy(), device=dev),
                        torch.as_tensor(a["earth"][:, :1].copy(), device=dev), return_comps=True,
                        precision=precision, peer_map=pmc)
            result["cyclic_array"] = pmc.finish().cpu().numpy()
            dist.barrier()
            pmc.close()
        queue.put((rank, result))
    except Exception as err:
        queue.put((rank, err))
        raise
    finally:
        dist.destroy_process_group()


WAR_SCALES = (1.001, 0.999, 1.002)  # observer scalings of the repeated evaluations (stay inside every cutoff)


def sharding_flags(dm, r_max):
    from zodipy_b200.spec import outside_flags

    return outside_flags(dm.spec, r_max)


def _run(world, backend, case_id, per_sample, precision, use_peer_map):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, case_id, per_sample, precision,
                                               use_peer_map, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r, v in results.items():
        assert not isinstance(v, Exception), f"rank {r}: {v!r}"
    return results


@pytest.mark.parametrize("case_id,per_sample", [("dirbe_25um_tod_straddle", True), ("planck18_857", False),
                                                ("rrm_60um", False)])
def test_two_processes_one_gpu_bitwise(case_id, per_sample):
    from helpers import golden_case
    from zodipy_b200 import engine

    case, a = golden_case(case_id)
    single = engine.DeviceModel(case["spec"], 0).evaluate(a["u"], a["obs"], a["earth"], return_comps=True)
    results = _run(2, "gloo", case_id, per_sample, "fp64", False)
    for r in range(2):
        np.testing.assert_array_equal(results[r]["gathered"], single)


def test_fused_peer_store_two_processes_one_gpu():
    """The fused gather on a 1-GPU box: two processes share cuda:0 (gloo for the set-up), each maps the
    other's buffers through CUDA IPC, the kernels store into both maps and the flag rendezvous - spinning
    kernels of two time-sliced processes - completes.  Same checks as the multi-GPU test (single evaluation,
    repeated evaluations into one PeerMap, block-cyclic shards with on-device directions)."""
    from helpers import golden_case
    from zodipy_b200 import engine

    case_id = "planck18_857"
    case, a = golden_case(case_id)
    dm1 = engine.DeviceModel(case["spec"], 0)
    single = dm1.evaluate(a["u"], a["obs"], a["earth"], return_comps=True)
    singles_scaled = [dm1.evaluate(a["u"], a["obs"] * s, a["earth"], return_comps=True,
                                   outside_flags=sharding_flags(dm1, dm1.max_observer_radius(a["obs"]) * s))
                      for s in WAR_SCALES]
    hp_single = dm1.evaluate_healpix(16, a["obs"][:, :1], a["earth"][:, :1], return_comps=True)
    results = _run(2, "gloo", case_id, False, "fp64", True)
    for r in range(2):
        np.testing.assert_array_equal(results[r]["fused"], single)
        for k in range(len(WAR_SCALES)):
            np.testing.assert_array_equal(results[r]["war"][k], singles_scaled[k])
        np.testing.assert_array_equal(results[r]["cyclic_healpix"], hp_single)
        np.testing.assert_allclose(results[r]["cyclic_array"], hp_single, rtol=1e-12)


@pytest.mark.parametrize("case_id,per_sample", [("dirbe_25um_tod_straddle", True), ("planck18_857", False)])
def test_fused_peer_store_equals_allgather(case_id, per_sample):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    from helpers import golden_case
    from zodipy_b200 import engine

    world = min(torch.cuda.device_count(), 4)
    case, a = golden_case(case_id)
    single = engine.DeviceModel(case["spec"], 0).evaluate(a["u"], a["obs"], a["earth"], return_comps=True)
    results = _run(world, "nccl", case_id, per_sample, "fp64", True)
    dm1 = engine.DeviceModel(case["spec"], 0)
    singles_scaled = [dm1.evaluate(a["u"], a["obs"] * s, a["earth"], return_comps=True,
                                   outside_flags=sharding_flags(dm1, dm1.max_observer_radius(a["obs"]) * s))
                      for s in WAR_SCALES]
    hp_single = engine.DeviceModel(case["spec"], 0).evaluate_healpix(
        16, a["obs"][:, :1], a["earth"][:, :1], return_comps=True)
    for r in range(world):
        np.testing.assert_array_equal(results[r]["gathered"], single)
        np.testing.assert_array_equal(results[r]["fused"], single)
        for k, scale in enumerate(WAR_SCALES):
            np.testing.assert_array_equal(results[r]["war"][k], singles_scaled[k])
        np.testing.assert_array_equal(results[r]["cyclic_healpix"], hp_single)
        # directions from the host pix2vec differ from the device routine in the last ulp
        np.testing.assert_allclose(results[r]["cyclic_array"], hp_single, rtol=1e-12)


def _worker_persistent(rank, world, port, backend, nside, name, x, unit, queue):
    """fp32 map large enough for the persistent-tile form of the packed kernel (peer stores), three evaluations
    back to back into one double-buffered PeerMap: the tile counters reset themselves between launches."""
    import torch
    import torch.distributed as dist

    import zodipy_b200 as zp
    from zodipy_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ZODI_X2_PERSIST="1")  # opt-in knob
    dev_index = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dm = zp.Model(zp.Quantity(x, unit), name=name, precision="fp32", device=dev_index).device_model
        npix = 12 * nside * nside
        out = {}
        for return_comps in (False, True):
            pm = sharding.PeerMap(npix, dm.ncomps if return_comps else 1, np.float32, dev_index, cyclic_block=4096)
            maps = []
            for scale in (1.0, 1.001, 1.0):
                dm.evaluate_healpix(nside, EARTH * scale, EARTH, return_comps=return_comps, precision="fp32",
                                    out_dtype=np.float32, peer_map=pm)
                maps.append(pm.finish().cpu().numpy().copy())
            assert not pm.timed_out()
            dist.barrier()
            pm.close()
            out[return_comps] = maps
        queue.put((rank, out))
    except Exception as err:
        queue.put((rank, err))
        raise
    finally:
        dist.destroy_process_group()


EARTH = np.array([[-0.3919640703], [0.9020953332], [0.0]])


@pytest.mark.parametrize("name,x,unit", [("planck18", 857.0, "GHz"), ("dirbe", 25.0, "um")])
def test_persistent_tiles_with_peer_stores(name, x, unit):
    """Sharded fp32 map through the persistent-tile packed kernel == the one-launch single-device map, bit for
    bit, on every rank and for repeated evaluations (2 processes on one GPU, or one per GPU when there are 2)."""
    import torch
    import torch.multiprocessing as mp

    import zodipy_b200 as zp

    nside = 512  # 1.57 M lines of sight per rank = 6144 tiles > the 1480 resident CTAs
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    dm1 = zp.Model(zp.Quantity(x, unit), name=name, precision="fp32").device_model
    singles = {rc: [dm1.evaluate_healpix(nside, EARTH * s, EARTH, return_comps=rc, precision="fp32",
                                         out_dtype=np.float32) for s in (1.0, 1.001)] for rc in (False, True)}
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_persistent, args=(r, 2, port, backend, nside, name, x, unit, queue))
             for r in range(2)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for r, v in results.items():
        assert not isinstance(v, Exception), f"rank {r}: {v!r}"
        for rc in (False, True):
            for got, want in zip(v[rc], (singles[rc][0], singles[rc][1], singles[rc][0])):
                np.testing.assert_array_equal(got.reshape(want.shape), want)
