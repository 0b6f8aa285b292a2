"""Multi-GPU execution: shard the time axis across ranks, gather for the writer.

The reference parallelises with ``multiprocessing.Pool`` over (time chunk x
grid-row chunk) tasks and exchanges results through the netCDF file on disk
(interp/main.py:84-153, :652-859).  Here one process drives one GPU
(``torch.distributed``, NCCL over NVLink on a B200 box, gloo in the CPU tests).
Time steps are independent units -- each step's availability group and system
are private to it -- so every rank takes a contiguous block of steps with
replicated coordinates and there is NO collective on the data path.  The only
exchange is the final gather of the f32 ``[T_rank, ny*nx]`` slabs to the writer
rank (HDF5 is single-writer without MPI-IO).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_steps, world_size, weights=None):
    """Contiguous blocks [b_r, b_{r+1}) of steps, one per rank.

    Without weights the split is ``np.linspace`` like misc.py:601-613
    (``ret_mp_idxs``).  With per-step weights (e.g. the cost model
    ``m_g^3 / steps_in_group + m_g * G``) the cumulative weight is balanced.
    """
    n_steps = int(n_steps)
    world_size = int(world_size)
    if weights is None:
        return np.linspace(0, n_steps, world_size + 1, endpoint=True, dtype=np.int64)
    w = np.asarray(weights, dtype=np.float64)
    assert w.shape == (n_steps,) and (w >= 0).all()
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(world_size + 1) / world_size
    b = np.searchsorted(cum, targets, side='left').astype(np.int64)
    b[0], b[-1] = 0, n_steps
    return np.maximum.accumulate(b)


def my_shard(n_steps, weights=None, group=None):
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    b = shard_bounds(n_steps, world, weights)
    return int(b[rank]), int(b[rank + 1])


def gather_slabs(slab, bounds, dst=0, group=None):
    """Gather ``[T_rank, ...]`` slabs (uneven T_rank allowed) onto rank ``dst``.

    Returns the assembled ``[T, ...]`` tensor on ``dst`` (on the slab's device)
    and ``None`` elsewhere.  Point-to-point sends: every byte crosses NVLink once
    and only towards the writer.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return slab
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    bounds = np.asarray(bounds)
    assert slab.shape[0] == int(bounds[rank + 1] - bounds[rank])
    if rank == dst:
        full = torch.empty((int(bounds[-1]),) + tuple(slab.shape[1:]), dtype=slab.dtype,
                           device=slab.device)
        full[int(bounds[rank]):int(bounds[rank + 1])] = slab
        reqs = []
        for r in range(world):
            if r == dst or bounds[r + 1] == bounds[r]:
                continue
            reqs.append(dist.irecv(full[int(bounds[r]):int(bounds[r + 1])], src=r, group=group))
        for q in reqs:
            q.wait()
        return full
    if slab.shape[0]:
        dist.send(slab.contiguous(), dst=dst, group=group)
    return None


def broadcast_inputs(arrays, src=0, group=None, device=None):
    """Broadcast a dict of NumPy arrays from ``src`` (coordinates, data, drift --
    MBs) so that only one rank has to read the inputs."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return arrays
    rank = dist.get_rank(group)
    meta = [{k: (v.shape, str(v.dtype)) for k, v in arrays.items()}] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src, group=group)
    out = {}
    for k, (shape, dtype) in meta[0].items():
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(arrays[k]))
        else:
            t = torch.empty(shape, dtype=getattr(torch, dtype) if hasattr(torch, dtype)
                            else torch.from_numpy(np.empty(0, dtype=dtype)).dtype)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=src, group=group)
        out[k] = t.cpu().numpy()
    return out


def run_time_sharded(compute_slab, n_steps, weights=None, dst=0, group=None):
    """``compute_slab(beg, end) -> {label: tensor[end-beg, ...]}`` on every rank;
    returns ``{label: tensor[n_steps, ...]}`` on ``dst`` and ``None`` elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(n_steps, world, weights)
    slabs = compute_slab(int(bounds[rank]), int(bounds[rank + 1]))
    out = {}
    for lab in sorted(slabs):
        out[lab] = gather_slabs(slabs[lab], bounds, dst=dst, group=group)
    return out if rank == dst else None
