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
    gathered = torch.empty((world,) + lead + (slot,), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, send.contiguous(), group=group)
    if n % world == 0:
        # (world, ..., slot) -> (..., world * slot) without a trim
        return gathered.movedim(0, -2).reshape(lead + (world * slot,))
    parts = [gathered[r, ..., : hi - lo] for r, (lo, hi) in enumerate(bounds)]
    assert parts[rank].shape[-1] == local.shape[-1]
    return torch.cat(parts, dim=-1)
