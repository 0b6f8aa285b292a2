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


def plan_tasks(time_bounds, row_bounds, world_size):
    """(time chunk x grid-row chunk) tasks of a job and their owner ranks -- the GPU
    counterpart of the reference's task list (interp/main.py:84-153 maps
    ``ret_mp_idxs`` chunks of BOTH axes over a process pool, :733-734, :805-811).

    One row chunk: contiguous blocks of time chunks per rank (each rank keeps its
    per-variogram caches warm).  Several row chunks (few steps, huge grid): the row
    chunk decides the owner, so that a rank keeps ONE grid geometry.
    Returns a list of (t_beg, t_end, r_beg, r_end, owner) in a deterministic order that
    every rank computes identically."""
    tb = [int(v) for v in time_bounds]
    rb = [int(v) for v in row_bounds]
    n_t, n_r = len(tb) - 1, len(rb) - 1
    world_size = int(world_size)
    tasks = []
    if n_r <= 1:
        own = shard_bounds(n_t, world_size)
        for i in range(n_t):
            owner = int(np.searchsorted(own, i, side='right') - 1)
            tasks.append((tb[i], tb[i + 1], rb[0], rb[-1], min(owner, world_size - 1)))
    else:
        for j in range(n_r):
            for i in range(n_t):
                tasks.append((tb[i], tb[i + 1], rb[j], rb[j + 1], j % world_size))
    return [t for t in tasks if t[1] > t[0] and t[3] > t[2]]


class StreamedGather:
    """Round-by-round exchange of finished task slabs with the writer rank.

    Every rank walks its own task list; in round j the writer receives the j-th slab of
    every other rank into a small ring of receive buffers (``depth`` slots of the largest
    slab), hands it to ``consume`` and reuses the slot -- the writer never holds more than
    ``depth`` foreign slabs, however long the job is (the reference exchanges results
    through locked writes into the netCDF file, interp/steps.py:895-954).  Senders keep a
    slab alive until the send after next has been posted.  Point-to-point ``isend`` /
    ``irecv``: NCCL over NVLink on a B200 box, gloo in the CPU tests."""

    def __init__(self, tasks, labels, writer=0, group=None, depth=2):
        self.tasks = list(tasks)
        self.labels = list(labels)
        self.writer = int(writer)
        self.group = group
        self.depth = int(depth)
        self.multi = dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.multi else 0
        self.world = dist.get_world_size(group) if self.multi else 1
        self.by_rank = [[t for t in self.tasks if t[4] == r] for r in range(self.world)]
        self.n_rounds = max((len(b) for b in self.by_rank), default=0)
        self._ring = None
        self._ring_k = 0
        self._inflight = []          # (work handles, tensors kept alive)
        self.bytes_received = 0

    def my_tasks(self):
        return self.by_rank[self.rank]

    @staticmethod
    def slab_shape(task, n_cols):
        return (task[1] - task[0], (task[3] - task[2]) * int(n_cols))

    def send(self, slabs, stats):
        """Non-writer: post the sends of one finished task ({label: tensor [T, cells]},
        {label: tensor [5, T] float64})."""
        works, keep = [], []
        for lab in self.labels:
            for t in (slabs[lab], stats[lab]):
                t = t.contiguous()
                works.append(dist.isend(t, dst=self.writer, group=self.group))
                keep.append(t)
        self._inflight.append((works, keep))
        while len(self._inflight) > 2:
            for w in self._inflight.pop(0)[0]:
                w.wait()

    def flush(self):
        for works, _ in self._inflight:
            for w in works:
                w.wait()
        self._inflight = []

    def receive_round(self, j, n_cols, dtype, device, consume):
        """Writer: receive the j-th task of every other rank, one at a time, and call
        ``consume(task, label, tensor [T, cells], stats tensor [5, T])`` for each label."""
        for r in range(self.world):
            if r == self.writer or j >= len(self.by_rank[r]):
                continue
            task = self.by_rank[r][j]
            shape = self.slab_shape(task, n_cols)
            if self._ring is None:
                big = max(self.slab_shape(t, n_cols)[0] * self.slab_shape(t, n_cols)[1]
                          for t in self.tasks)
                self._ring = [torch.empty(big, dtype=dtype, device=device)
                              for _ in range(self.depth)]
            for lab in self.labels:
                buf = self._ring[self._ring_k % self.depth][:shape[0] * shape[1]].view(shape)
                self._ring_k += 1
                st = torch.empty((5, shape[0]), dtype=torch.float64, device=device)
                w1 = dist.irecv(buf, src=r, group=self.group)
                w2 = dist.irecv(st, src=r, group=self.group)
                w1.wait()
                w2.wait()
                self.bytes_received += buf.numel() * buf.element_size()
                # the slot is reused ``depth`` receives later: consume() may leave an
                # asynchronous copy of ``buf`` in flight for that long
                consume(task, lab, buf, st)


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
