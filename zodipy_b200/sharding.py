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

    Every rank allocates TWO (rows, n_total) buffers with the library (plain ``cudaMalloc`` so that they
    can be exported) plus a small flag array, the IPC handles are exchanged once with
    ``all_gather_object``, and each rank maps every peer's buffers.
    ``DeviceModel.evaluate(..., peer_map=pm)`` makes the compute kernel store this rank's slice into
    the current buffer of ALL ranks (fused all-gather); ``finish()`` enqueues the completion rendezvous
    (one tiny kernel, ``zodi_peer_rendezvous``: publish this rank's epoch to every peer, wait for
    theirs) behind it and returns the assembled map, valid in stream order on every rank.

    Double buffering closes the write-after-read hazard of repeated evaluations: evaluation i + 1
    writes the OTHER buffer, and evaluation i + 2 - which reuses the buffer of evaluation i - can only
    start on any rank after that rank passed rendezvous i + 1, i.e. after every peer launched its
    kernel i + 1, which is stream-ordered behind that peer's reads of map i.  So: read the returned map
    on the stream that calls ``evaluate`` (or make that stream wait for the reader) and it is safe to
    keep evaluating while peers are still reading the previous map.  One process per GPU, ranks of one
    NVSwitch box.
    """

    N_BUFFERS = 2

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
        self.world, self.rank = world, rank
        # contiguous np.array_split shards by default; block-cyclic (load-balanced) on request
        self.cyclic = (int(cyclic_block), world, rank) if cyclic_block else None
        self.offset = 0 if self.cyclic else split_bounds(self.n_total, world)[rank][0]
        nbytes = self.rows * self.n_total * self.dtype.itemsize
        sizes = [nbytes] * self.N_BUFFERS + [4 * (_cabi.MAX_PEERS + 1)]  # maps + flag array (zeroed by alloc)
        self._own, mine = [], []
        for size in sizes:
            own = C.c_void_p()
            handle = (C.c_uint8 * _cabi.IPC_HANDLE_BYTES)()
            _cabi.check(self._lib.zodi_peer_buffer_alloc(self.device_index, size, C.byref(own), handle))
            self._own.append(own.value)
            mine.append(bytes(handle))
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        self._opened = []
        mapped = [[] for _ in sizes]  # [buffer][rank] -> pointer in this process
        for r, hs in enumerate(handles):
            for b, h in enumerate(hs):
                if r == rank:
                    mapped[b].append(self._own[b])
                    continue
                buf = (C.c_uint8 * _cabi.IPC_HANDLE_BYTES).from_buffer_copy(h)
                ptr = C.c_void_p()
                _cabi.check(self._lib.zodi_peer_buffer_open(self.device_index, buf, C.byref(ptr)))
                mapped[b].append(ptr.value)
                self._opened.append(ptr.value)
        self._map_pointers = mapped[:self.N_BUFFERS]
        self._flag_pointers = (C.c_void_p * world)(*mapped[self.N_BUFFERS])
        # torch views of the OWN buffers (no copy)
        tdtype = torch.float32 if self.dtype == np.float32 else torch.float64
        shape = (self.rows, self.n_total) if self.rows > 1 else (self.n_total,)
        self._views = []
        for b in range(self.N_BUFFERS):
            iface = {"shape": shape, "typestr": "<f4" if self.dtype == np.float32 else "<f8",
                     "data": (self._own[b], False), "version": 2}
            holder = type("_Buf", (), {"__cuda_array_interface__": iface})()
            view = torch.as_tensor(holder, device=torch.device("cuda", self.device_index))
            assert view.dtype == tdtype and view.data_ptr() == self._own[b]
            self._views.append(view)
        flag_iface = {"shape": (_cabi.MAX_PEERS + 1,), "typestr": "<i4", "data": (self._own[-1], False), "version": 2}
        self._flags_view = torch.as_tensor(type("_Buf", (), {"__cuda_array_interface__": flag_iface})(),
                                           device=torch.device("cuda", self.device_index))
        self._cur = 0      # buffer the next evaluation writes
        self._epoch = 0
        self.tensor = self._views[0]
        # every rank must have mapped (and the owner zeroed) the flag arrays before the first rendezvous
        torch.cuda.synchronize(self.device_index)
        dist.barrier(group=group)

    @property
    def pointers(self):
        """Every rank's CURRENT map buffer as mapped in this process (the kernel's peer_out[])."""
        return self._map_pointers[self._cur]

    def finish(self):
        """Stream-ordered completion: enqueues the rendezvous kernel behind the integrator kernel and
        returns this rank's assembled map (valid for work enqueued on the same stream afterwards).  The
        next evaluation writes the other buffer."""
        import torch

        self._epoch += 1
        stream = torch.cuda.current_stream(self.device_index).cuda_stream
        from . import _cabi

        _cabi.check(self._lib.zodi_peer_rendezvous(self.device_index, self._flag_pointers, self.world, self.rank,
                                                   self._epoch & 0xFFFFFFFF, stream))
        self.tensor = self._views[self._cur]
        self._cur = (self._cur + 1) % self.N_BUFFERS
        return self.tensor

    def timed_out(self) -> bool:
        """True if a rendezvous gave up waiting for a peer (synchronises the device)."""
        from . import _cabi

        return bool(self._flags_view[_cabi.MAX_PEERS].item())

    def close(self):
        import torch

        torch.cuda.synchronize(self.device_index)
        for ptr in self._opened:
            self._lib.zodi_peer_buffer_close(self.device_index, ptr)
        self._opened = []
        if self._own:
            self.tensor = None
            self._views = []
            self._flags_view = None
            for ptr in self._own:
                self._lib.zodi_peer_buffer_free(self.device_index, ptr)
            self._own = []
