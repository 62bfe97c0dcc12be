"""Sharding of lines of sight over the GPUs of one box (one process per GPU).

Replaces the reference's only parallel strategy, the ``multiprocessing`` fork pool of
``zodipy/model.py:182-198``: contiguous ``np.array_split`` chunks (first ``N mod P`` ranks get one
extra element), per-sample observer/Earth arrays split the same way, instantaneous positions and
model parameters replicated.  Each rank evaluates its slice; the map is assembled on every rank
by an all-gather (NCCL over NVLink in production, gloo in the CPU tests).

The early-out flags of ``get_sphere_intersection`` are a GLOBAL ``.any()`` over all observers
(``zodipy/line_of_sight.py:72-73``): each rank reduces its local largest observer
distance, a MAX all-reduce makes it global, and every rank derives identical flags - so the sharded
result is bit-identical to the single-GPU one.
"""
from __future__ import annotations

import numpy as np


def split_bounds(n: int, parts: int) -> list[tuple[int, int]]:
    """[start, stop) of each ``np.array_split`` chunk (zodipy/model.py:184)."""
    base, extra = divmod(n, parts)
    bounds, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


CYCLIC_BLOCK = 1 << 16


def cyclic_count(n: int, parts: int, rank: int, block: int = CYCLIC_BLOCK) -> int:
    """Number of elements rank ``rank`` owns under the block-cyclic layout (blocks rank,
    rank + parts, ... of ``block`` elements; only the globally last block may be short)."""
    nblocks = -(-n // block)
    mine = len(range(rank, nblocks, parts))
    if mine == 0:
        return 0
    last_global = rank + (mine - 1) * parts
    short = block - (n - last_global * block) if last_global == nblocks - 1 else 0
    return mine * block - max(short, 0)


def cyclic_indices(n: int, parts: int, rank: int, block: int = CYCLIC_BLOCK) -> np.ndarray:
    """Global indices g(j), j = 0..count-1, of rank ``rank``'s block-cyclic shard (the same map the
    kernels apply: ``((j // block) * parts + rank) * block + j % block``)."""
    j = np.arange(cyclic_count(n, parts, rank, block), dtype=np.int64)
    return ((j // block) * parts + rank) * block + j % block


def padded_count(n: int, parts: int) -> int:
    """Equal per-rank slot length used by the all-gather (ceil(n / parts))."""
    return -(-n // parts)


def global_max_radius(local_r_max: float, group=None) -> float:
    """MAX all-reduce of the per-rank largest observer distance."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_r_max
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([local_r_max], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allgather_map(local, n: int, group=None):
    """Assemble the (..., n) map from per-rank (..., n_r) slices (torch tensors).

    Slices are padded to ``padded_count`` so one ``all_gather_into_tensor`` suffices; the padding
    is trimmed afterwards (first ``n mod P`` ranks hold one more element than the others).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    bounds = split_bounds(n, world)
    slot = padded_count(n, world)
    lead = tuple(local.shape[:-1])
    send = local
    if local.shape[-1] != slot:
        send = torch.zeros(lead + (slot,), dtype=local.dtype, device=local.device)
        send[..., : local.shape[-1]] = local
    # flat 1-D buffers: accepted by both NCCL and gloo (rank r's slot is gathered[r])
    send = send.contiguous().reshape(-1)
    on_cuda = send.is_cuda
    if on_cuda and dist.get_backend(group) == "gloo":
        send = send.cpu()  # gloo moves host memory (CPU tests / single-GPU multi-process tests)
    flat = torch.empty(world * send.numel(), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(flat, send, group=group)
    if on_cuda and not flat.is_cuda:
        flat = flat.to(local.device)
    gathered = flat.view((world,) + lead + (slot,))
    if n % world == 0:
        # (world, ..., slot) -> (..., world * slot) without a trim
        return gathered.movedim(0, -2).reshape(lead + (world * slot,))
    parts = [gathered[r, ..., : hi - lo] for r, (lo, hi) in enumerate(bounds)]
    assert parts[rank].shape[-1] == local.shape[-1]
    return torch.cat(parts, dim=-1)


def evaluate_sharded(evaluate_fn, max_radius_fn, flags_fn, u_local, obs_local, earth_local, n_total,
                     *, obs_per_sample: bool, group=None, gather: bool = True, **kwargs):
    """One rank's part of a sharded evaluation.

    ``u_local`` is this rank's contiguous slice (``split_bounds`` rule) of the (3, n_total)
    directions; ``obs_local`` / ``earth_local`` are either the rank's slice of per-sample positions
    (``obs_per_sample``) or the replicated single position.  Steps: local max observer distance ->
    MAX all-reduce -> identical global early-out flags on every rank (quirk Q1) -> local
    evaluation -> all-gather of the map.  ``evaluate_fn(u, obs, earth, outside_flags=..., **kwargs)``
    and ``max_radius_fn(obs)`` / ``flags_fn(r_max)`` are normally bound methods of a
    :class:`zodipy_b200.engine.DeviceModel`; they are parameters so the plumbing can be tested
    on CPU with the gloo backend.
    """
    r_local = max_radius_fn(obs_local)
    r_global = global_max_radius(r_local, group) if obs_per_sample else r_local
    flags = flags_fn(r_global)
    local = evaluate_fn(u_local, obs_local, earth_local, outside_flags=flags, **kwargs)
    if not gather:
        return local
    return allgather_map(local, n_total, group)


class PeerMap:
    """Full-map output buffers of all ranks, mapped into this process (CUDA IPC over NVLink).

    Every rank allocates one (rows, n_total) buffer with the library (plain ``cudaMalloc`` so
    that it can be exported), the IPC handles are exchanged once with ``all_gather_object``, and
    each rank maps every peer's buffer.  ``DeviceModel.evaluate(..., peer_map=pm)`` then makes the
    compute kernel store this rank's slice into all of them (fused all-gather); ``finish()``
    is the stream-ordered rendezvous after which ``pm.tensor`` holds the complete map on every
    rank.  One process per GPU, ranks of one NVSwitch box.
    """

    def __init__(self, n_total: int, rows: int, dtype, device_index: int, group=None,
                 cyclic_block: int = 0):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _cabi

        self._lib = _cabi.load()
        self.n_total, self.rows = int(n_total), int(rows)
        self.dtype = np.dtype(dtype)
        self.device_index = int(device_index)
        self.group = group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > _cabi.MAX_PEERS:
            raise ValueError(f"at most {_cabi.MAX_PEERS} peers")
        # contiguous np.array_split shards by default; block-cyclic (load-balanced) on request
        self.cyclic = (int(cyclic_block), world, rank) if cyclic_block else None
        self.offset = 0 if self.cyclic else split_bounds(self.n_total, world)[rank][0]
        nbytes = self.rows * self.n_total * self.dtype.itemsize
        own = C.c_void_p()
        handle = (C.c_uint8 * _cabi.IPC_HANDLE_BYTES)()
        _cabi.check(self._lib.zodi_peer_buffer_alloc(self.device_index, nbytes, C.byref(own), handle))
        self._own = own.value
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.pointers, self._opened = [], []
        for r, h in enumerate(handles):
            if r == rank:
                self.pointers.append(self._own)
                continue
            buf = (C.c_uint8 * _cabi.IPC_HANDLE_BYTES).from_buffer_copy(h)
            ptr = C.c_void_p()
            _cabi.check(self._lib.zodi_peer_buffer_open(self.device_index, buf, C.byref(ptr)))
            self.pointers.append(ptr.value)
            self._opened.append(ptr.value)
        # torch view of the OWN buffer (no copy)
        tdtype = torch.float32 if self.dtype == np.float32 else torch.float64
        shape = (self.rows, self.n_total) if self.rows > 1 else (self.n_total,)
        iface = {"shape": shape, "typestr": "<f4" if self.dtype == np.float32 else "<f8",
                 "data": (self._own, False), "version": 2}
        holder = type("_Buf", (), {"__cuda_array_interface__": iface})()
        self.tensor = torch.as_tensor(holder, device=torch.device("cuda", self.device_index))
        assert self.tensor.dtype == tdtype and self.tensor.data_ptr() == self._own
        self._token = torch.zeros(1, dtype=torch.float32, device=self.tensor.device)

    def finish(self):
        """Stream-ordered rendezvous: returns once every rank's kernel (and therefore all of its
        peer stores) has completed.  A 4-byte NCCL all-reduce enqueued behind the kernel."""
        import torch.distributed as dist

        dist.all_reduce(self._token, group=self.group)
        return self.tensor

    def close(self):
        import torch

        torch.cuda.synchronize(self.device_index)
        for ptr in self._opened:
            self._lib.zodi_peer_buffer_close(self.device_index, ptr)
        self._opened = []
        if self._own:
            self.tensor = None
            self._lib.zodi_peer_buffer_free(self.device_index, self._own)
            self._own = None
