"""GPU engine for one (time chunk x grid-row chunk) of the interpolation.

This is the B200 replacement of the compute half of
``SpInterpSteps.interpolate_subset`` (reference interp/steps.py:478-877).  The
host side (this file) keeps the reference's *index* logic -- cell subsetting,
availability groups, variogram clusters, per-step flags -- and turns it into
dense descriptor arrays; all floating-point work runs in the sm_100a kernels of
``csrc/`` through the C-ABI (include/spx_b200.h):

  assemble -> LU factor -> solve          (one system per group x variogram)
  fused variogram-fill + DMMA contraction  (dual form: Z = C . RHS^T)
  fused IDW (same contraction with d**-p weights and a sum-of-weights row)
  nearest neighbour index / gather, constant rows

torch is used for device memory, streams and host<->device copies only.
There is no CPU fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import os
import time
import types

import numpy as np
import torch

from . import _lib
from .vgs import check_full_nuggetness, vg_abs_bound

_I32 = torch.int32
_I64 = torch.int64
_F64 = torch.float64


def _pad_up(n, m):
    return ((int(n) + m - 1) // m) * m


def availability_groups(avail):
    """Distinct rows of the [T, N] availability matrix in first-occurrence
    order (interp/grps.py:57-101 groups steps by the set of non-NaN stations).

    Returns grp_of_step [T] int32 and grp_mask [n_grps, N] bool.
    """
    T, N = avail.shape
    packed = np.packbits(avail, axis=1)
    keys = np.ascontiguousarray(packed).view(np.dtype((np.void, packed.shape[1]))).ravel()
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')          # unique-id -> rank by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    grp_of_step = rank[inv.ravel()].astype(np.int32)
    grp_mask = avail[first[order]]
    return grp_of_step, grp_mask


def station_clusters(sx, sy, R, max_size):
    """Connected components of the graph "stations closer than R" as labels 0..n_comp-1 per
    station, or None when a component has more than ``max_size`` stations (or the graph is
    dense).  NumPy only: importing scipy here would cost the first chunk of a job 0.4 s."""
    sx = np.asarray(sx, dtype=np.float64)
    sy = np.asarray(sy, dtype=np.float64)
    n = sx.size
    # a little beyond R: pairs at exactly the range have vg == F anyway (their C entry is 0)
    r2 = (R * (1.0 + 1e-12) + 1e-9) ** 2
    ei, ej = [], []
    n_pairs = 0
    for b0 in range(0, n, 2048):                      # row blocks of the distance matrix
        b1 = min(n, b0 + 2048)
        d2 = (sx[b0:b1, None] - sx[None, :]) ** 2 + (sy[b0:b1, None] - sy[None, :]) ** 2
        ii, jj = np.nonzero(d2 < r2)
        keep = jj > ii + b0
        ei.append(ii[keep] + b0)
        ej.append(jj[keep])
        n_pairs += int(keep.sum())
        if n_pairs > 8 * n:
            return None                               # dense graph: no small clusters
    ei, ej = np.concatenate(ei), np.concatenate(ej)
    # min-label propagation converges within the diameter of the largest component, so more
    # than max_size rounds means a component that is too large
    lab = np.arange(n, dtype=np.int64)
    for _ in range(int(max_size) + 1):
        new = lab.copy()
        np.minimum.at(new, ei, lab[ej])
        np.minimum.at(new, ej, lab[ei])
        if np.array_equal(new, lab):
            break
        lab = new
    else:
        return None
    _, lab = np.unique(lab, return_inverse=True)
    if np.bincount(lab).max() > max_size:
        return None
    return lab


class _LazyCtx(dict):
    """Per-chunk context; entries of ``lazy`` (byte masks, NaN -> 0 data, ...) are built on
    first use -- the common path needs none of them."""
    lazy = None

    def __missing__(self, key):
        fn = (self.lazy or {}).get(key)
        if fn is None:
            raise KeyError(key)
        val = fn()
        self[key] = val
        return val


class PendingChunk:
    """A chunk whose kernels are queued on the stream.  ``result()`` runs the
    deferred health checks (reading back a few flags copied to pinned memory
    before the contraction was queued), applies the rare fix-ups and returns the
    fields.  Submitting chunk i+1 before asking for the result of chunk i hides
    the host-side preparation behind the GPU work of the previous chunk."""

    def __init__(self, engine, flds, problem_steps, deferred, stats):
        self.engine = engine
        self.flds = flds
        self.problem_steps = problem_steps
        self.deferred = deferred
        self.stats = stats
        # optional output stage (interp/steps.py:907-912, interp/main.py:474-525):
        # round to `round_decimals` places and / or reduce per-step field statistics
        self.round_decimals = None
        self.want_field_stats = False
        self.field_stats = None        # {label: ndarray[5, T]} min, mean, max, std, count
        self._d_field_stats = None
        self.done_event = torch.cuda.Event()
        self.done_event.record(torch.cuda.current_stream(engine.device))

    def _output_stage(self, labels=None):
        eng = self.engine
        if labels is None:
            if self._d_field_stats is not None or (
                    self.round_decimals is None and not self.want_field_stats):
                return
            self._d_field_stats = {}
        dec = -1 if self.round_decimals is None else int(self.round_decimals)
        for lab, t in self.flds.items():
            if labels is not None and lab not in labels:
                continue
            n_rows, row_len = t.shape
            st = torch.empty((5, n_rows), dtype=_F64, device=eng.device)
            for r0 in range(0, n_rows, 65535):
                r1 = min(n_rows, r0 + 65535)
                ws = torch.empty(max(1, eng.lib.spx_round_stats_workspace(r1 - r0, row_len)),
                                 dtype=torch.uint8, device=eng.device)
                sub = torch.empty((5, r1 - r0), dtype=_F64, device=eng.device)
                _lib.check(eng.lib.spx_round_stats_dev(
                    C.c_void_p(t[r0:r1].data_ptr()), int(t.dtype == _F64), r1 - r0, row_len,
                    t.stride(0), dec, C.c_void_p(sub.data_ptr()), C.c_void_p(ws.data_ptr()),
                    eng._stream()), 'round_stats')
                eng._count('launches', 2)
                st[:, r0:r1] = sub
            self._d_field_stats[lab] = st

    def start_packed(self):
        """Queue the writer's output stage and the compact download of every field; returns a
        handle for ``finish_packed``.  Rounded float32 fields that are large enough go through
        ONE pass over the unrounded field (np.round + per-step statistics + delta encoding,
        spx_dpack_field_dev with SPX_DPACK_ROUND; the device copy stays unrounded) and cross
        PCIe in the delta transport form; everything else takes the separate rounding /
        statistics kernel and comes back as an array.  Nothing here waits for the GPU, so the
        next chunk can be submitted before ``finish_packed`` is called."""
        eng = self.engine
        with torch.cuda.device(eng.device):
            eng._begin_call()
            for fn in self.deferred:
                fn()
            self.deferred = []
            tickets = {}
            if self._d_field_stats is None and self.round_decimals is not None:
                dec = int(self.round_decimals)
                self._d_field_stats = {}
                rest = []
                for lab, t in self.flds.items():
                    dl = eng._packed_downloader(t, dec, depth=max(2, len(self.flds)))
                    if dl is None or dl.codec != 'delta':
                        rest.append(lab)
                        continue
                    st = (torch.empty((5, t.shape[0]), dtype=_F64, device=eng.device)
                          if self.want_field_stats else None)
                    tickets[lab] = (dl, dl.start(t, dec, round_here=True, stats=st,
                                                 write_back=False))
                    eng._count('launches', 1 if st is None else 2)
                    if st is not None:
                        self._d_field_stats[lab] = st
                if rest:
                    self._output_stage(labels=rest)
                    for lab in rest:
                        dl = eng._packed_downloader(self.flds[lab], dec,
                                                    depth=max(2, len(self.flds)))
                        if dl is not None:
                            tickets[lab] = (dl, dl.start(self.flds[lab], dec))
                if not self.want_field_stats:
                    self._d_field_stats = None
            else:
                self._output_stage()
                for lab, t in self.flds.items():
                    dl = eng._packed_downloader(t, self.round_decimals,
                                                depth=max(2, len(self.flds)))
                    if dl is not None:
                        tickets[lab] = (dl, dl.start(t, self.round_decimals))
            hs = ev = None
            if self._d_field_stats and self.field_stats is None:
                hs = {lab: eng._fetch_async(t) for lab, t in self._d_field_stats.items()}
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(eng.device))
            return dict(tickets=tickets, hs=hs, ev=ev)

    def finish_packed(self, handle):
        """({label: transfer.DeltaField / PackedField over a downloader slot -- the consumer
        calls ``.release()`` -- or ndarray}, problem_steps)."""
        eng = self.engine
        with torch.cuda.device(eng.device):
            if handle['hs'] is not None:
                handle['ev'].synchronize()
                self.field_stats = {lab: h.numpy().copy() for lab, h in handle['hs'].items()}
            out = {}
            tickets = handle['tickets']
            for lab, t in self.flds.items():
                if lab in tickets:
                    dl, tk = tickets[lab]
                    pf = dl.wait(tk)
                    pf.release = (lambda dl=dl, tk=tk: dl.release(tk))
                    out[lab] = pf
                else:
                    out[lab] = t.cpu().numpy()
            return out, self.problem_steps

    def result(self, to_host=True):
        eng = self.engine
        if to_host == 'packed':
            return self.finish_packed(self.start_packed())
        with torch.cuda.device(eng.device):
            eng._begin_call()
            n0 = eng.total_launches
            for fn in self.deferred:
                fn()
            self.deferred = []
            self._output_stage()
            if self._d_field_stats is not None and self.field_stats is None:
                hs = {lab: eng._fetch_async(t) for lab, t in self._d_field_stats.items()}
                torch.cuda.current_stream(eng.device).synchronize()
                self.field_stats = {lab: h.numpy().copy() for lab, h in hs.items()}
            if eng.total_launches != n0:
                # fix-up kernels were queued: the fields are final only after them
                self.done_event = torch.cuda.Event()
                self.done_event.record(torch.cuda.current_stream(eng.device))
            if not to_host:
                return self.flds, self.problem_steps
            torch.cuda.current_stream(eng.device).synchronize()
            out = {}
            for lab, t in self.flds.items():
                dl = eng._packed_downloader(t, self.round_decimals)
                # rounded f32 fields cross PCIe as 16-bit codes (transfer.py), bit-exact
                out[lab] = (dl.download(t, self.round_decimals) if dl is not None
                            else t.cpu().numpy())
            return out, self.problem_steps


class _TracedLib:
    """Wraps the ctypes library so that every ``spx_*_dev`` entry point is bracketed by
    CUDA events on the compute stream (ChunkEngine.trace_launches).  A measuring aid:
    bench.py uses it for the per-kernel breakdown of a step."""

    def __init__(self, eng, lib):
        self._eng = eng
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.endswith('_dev') or name == 'spx_upload_dev':
            return fn
        eng = self._eng

        def traced(*a):
            if not eng.trace_launches:
                return fn(*a)
            st = torch.cuda.current_stream(eng.device)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = fn(*a)
            e1.record(st)
            eng.trace.append((name, e0, e1))
            return rc
        return traced


class ChunkEngine:
    """Holds the device and tunables; ``interp_chunk`` is re-entrant."""

    def __init__(self, device=None, work_limit_bytes=4 << 30, aux_limit_bytes=8 << 30,
                 lambda_tol=1e-7):
        _lib.require_gpu()
        if not torch.cuda.is_available():
            raise _lib.SpxError('torch sees no CUDA device (no CPU fallback)')
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None
                                   else int(device))
        self.lib = _TracedLib(self, _lib.load())
        self.trace_launches = False
        self.trace = []           # (entry point, start event, end event)
        self.work_limit = int(work_limit_bytes)
        self.aux_limit = int(aux_limit_bytes)
        self.lambda_tol = float(lambda_tol)
        # downdated solves (spx_krige_downdate_dev) when a variogram is shared by at
        # least this many availability groups
        self.downdate = True
        self.downdate_min_systems = 4
        # all-downdated chunks are planned by the native host planner (csrc/spx_plan.cu)
        # instead of the NumPy index logic of _krige / _solve_downdate
        self.native_plan = True
        # ... and, when the whole chunk shares one variogram, submitted by ONE native call
        # (csrc/spx_chunk.cu) whose solve phase runs on a second stream underneath the
        # estimate kernel of the previous chunk
        self.native_submit = os.environ.get('SPX_NATIVE_SUBMIT', '1') != '0'
        self.fast_slots = 4
        self.solve_stream = False   # solve phase on its own stream (no gain)
        self._fast_jobs = {}
        self._fast_prof = {}      # (job id, slot) -> (kernel, bound, work) awaiting its time
        # full-system inverses are reused across chunks with the same stations and
        # variogram (they do not depend on the data)
        self.multivg = True      # per-row-variogram estimator for variogram series
        # local estimator for compactly supported variograms (Nug + Sph / Lin): used
        # when a cell has on average at most this many stations within the range
        self.local_support = True
        self.local_max_near = 24.0
        self.local_tiles = True     # shared-memory staging of the coefficient slices
        self._local_cache = {}
        self._geom_cache = {}
        self._token_cache = {}    # id(array) -> (weakref, sampled fingerprint, content token)
        self.pinv_flagged = True  # np.linalg.pinv semantics for untrustworthy OK/EDK systems
        self.ginv_cache = True
        self.ginv_cache_size = 8
        self._ginv_cache = {}
        self.stats = {}
        # bench hook: CUDA events around every estimate-contraction launch
        self.profile_gemm = False
        self.gemm_events = []
        self.kernel_events = []   # (kernel, bound, work [flop | bytes], start, end)
        self.sync_timing = False
        self.timing = {}
        self.total_launches = 0
        self.h2d_bytes = 0           # bytes uploaded by the engine (host -> device)
        self.h2d_stream = torch.cuda.Stream(self.device)
        self._h2d_handle = C.c_void_p(self.h2d_stream.cuda_stream)
        self._h2d_dirty = False
        self._main = None          # compute stream, cached per public call
        self._main_handle = None
        self._arena = None         # per-chunk upload arena (see _arena_take)
        self._arena_off = 0
        self._arenas = [None] * self._N_ARENAS
        self._arena_events = [None] * self._N_ARENAS
        self._arena_k = 0
        self._const_cache = {}
        # rounded float32 fields are downloaded as 16-bit codes and decoded on the host
        self.packed_download = True
        self._dl = None
        self._fast_jobs_none = set()   # job keys without inverse for which the sparse form was ruled out
        self.solve_ms = []             # solve-phase times of profiled native submits (profile_gemm)
        self.transport = None          # codec of the packed download (None: transfer.default_codec())
        # ordinary kriging with a compact variogram whose stations form small clusters:
        # solve the block-diagonal covariance form instead of the downdated dense systems
        self.sparse_solve = os.environ.get('SPX_SPARSE_SOLVE', '1') != '0'
        self.threaded_upload = True
        self._uploader = None

    # ------------------------------------------------------------ helpers
    _TORCH_OF = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                 np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                 np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool}
    _ARENA_MAX_ITEM = 32 << 20

    def _main_stream(self):
        if self._main is None:
            self._main = torch.cuda.current_stream(self.device)
            self._main_handle = C.c_void_p(self._main.cuda_stream)
        return self._main

    def _begin_call(self):
        """Refresh the cached compute stream (public entry points call this)."""
        self._main = None
        self._main_stream()

    _N_ARENAS = 4

    def _arena_next_chunk(self):
        """Switch to the next upload arena of a small ring (a fresh torch allocation per
        chunk costs a cudaMalloc: the previous arenas are still referenced by queued
        kernels).  An arena is reused only after the chunk that used it last has
        finished on the compute stream."""
        if self._arena is not None:                 # close the arena of the previous chunk
            ev = self._arena_events[self._arena_k]
            if ev is None:
                ev = self._arena_events[self._arena_k] = torch.cuda.Event()
            ev.record(self._main_stream())
        self._arena_k = (self._arena_k + 1) % self._N_ARENAS
        ev = self._arena_events[self._arena_k]
        if ev is not None:
            ev.synchronize()
        self._arena = self._arenas[self._arena_k]
        self._arena_off = 0

    def _arena_take(self, nbytes):
        """Device bytes for one upload, carved from the chunk's arena (allocated in the
        upload stream's pool, marked as used by the compute stream).  Requests that do
        not fit get a one-off allocation."""
        need = (nbytes + 255) & ~255
        if self._arena is None:
            with torch.cuda.stream(self.h2d_stream):
                self._arena = torch.empty(max(16 << 20, 2 * need), dtype=torch.uint8,
                                          device=self.device)
            self._arena.record_stream(self._main_stream())
            self._arenas[self._arena_k] = self._arena
            self._arena_off = 0
        if self._arena_off + need > self._arena.numel():
            with torch.cuda.stream(self.h2d_stream):
                t = torch.empty(need, dtype=torch.uint8, device=self.device)
            t.record_stream(self._main_stream())
            return t[:nbytes]
        o = self._arena_off
        self._arena_off += need
        return self._arena[o:o + nbytes]

    def _dev_threaded(self, arr):
        """Like _dev for one large array, but the (blocking) copy call is issued by a
        helper thread; returns (device view, future or None).  The caller must wait for
        the future before anything is launched that reads the view."""
        tdt = self._TORCH_OF.get(arr.dtype)
        if (not self.threaded_upload or tdt is None or arr.nbytes < (1 << 20)
                or arr.nbytes > self._ARENA_MAX_ITEM or not arr.flags.c_contiguous):
            return self._dev(arr), None
        if self._uploader is None:
            import concurrent.futures
            dev_index = self.device.index
            self._uploader = concurrent.futures.ThreadPoolExecutor(
                max_workers=1, thread_name_prefix='spx-upload',
                initializer=lambda: torch.cuda.set_device(dev_index))
        d = self._arena_take(arr.nbytes)
        dst, src, nbytes, stream = d.data_ptr(), arr.ctypes.data, arr.nbytes, self._h2d_handle

        def job():
            _lib.check(self.lib.spx_upload_dev(C.c_void_p(dst), C.c_void_p(src), nbytes, stream),
                       'upload')
        fut = self._uploader.submit(job)
        self._h2d_dirty = True
        self.h2d_bytes += arr.nbytes
        return d.view(tdt).view(arr.shape), fut

    def _array_token(self, arr):
        """Content token of a per-job input array (cell coordinates, cell mask): the keys
        of the cross-chunk caches.  The 16-byte content hash is computed once per array
        OBJECT; later chunks that pass the same object only pay for a strided sample of
        ~256 elements (the whole array when it has <= 4096), which detects in-place
        edits (shifted / rescaled / re-projected coordinates) and sends them through a
        full rehash.  Distinct objects with equal content share their token."""
        import hashlib
        import weakref
        a = np.asarray(arr)
        n = a.size
        flat = a.reshape(-1) if a.flags.c_contiguous else a.ravel()
        step = 1 if n <= 4096 else n // 256
        fp = (a.shape, a.dtype.str, hash(flat[::step].tobytes()),
              hash(flat[-1:].tobytes()))
        ent = self._token_cache.get(id(arr))
        if ent is not None and ent[0]() is arr and ent[1] == fp:
            return ent[2]
        tok = (a.shape, a.dtype.str,
               hashlib.blake2b(np.ascontiguousarray(a).view(np.uint8).reshape(-1),
                               digest_size=16).digest())
        try:
            ref = weakref.ref(arr)
        except TypeError:                   # not weak-referenceable (e.g. a list): no caching
            return tok
        if len(self._token_cache) >= 64:
            for k_ in [k_ for k_, v in self._token_cache.items() if v[0]() is None]:
                del self._token_cache[k_]
            while len(self._token_cache) >= 64:
                self._token_cache.pop(next(iter(self._token_cache)))
        self._token_cache[id(arr)] = (ref, fp, tok)
        return tok

    def _dev_const(self, arr):
        """Device copy of a small per-job constant (station coordinates): cached by
        content, so that the chunks of a job upload it once."""
        arr = np.ascontiguousarray(arr)
        key = (arr.dtype.str, arr.shape, arr.tobytes())
        hit = self._const_cache.get(key)
        if hit is None:
            self._sync_uploads()
            self.h2d_bytes += arr.nbytes
            hit = torch.from_numpy(arr.copy()).to(self.device)
            while len(self._const_cache) >= 16:
                self._const_cache.pop(next(iter(self._const_cache)))
            self._const_cache[key] = hit
        return hit

    def _dev_keep(self, arr):
        """Device copy of an array that OUTLIVES the chunk (cached geometry, bin tables):
        its own allocation -- the per-chunk upload arenas are recycled after four
        chunks."""
        arr = np.ascontiguousarray(arr)
        self._sync_uploads()
        self.h2d_bytes += arr.nbytes
        return torch.from_numpy(arr.copy()).to(self.device)

    def _dev(self, arr, dtype=None):
        """Host -> device copy on a dedicated upload stream.  A pageable-memory
        cudaMemcpyAsync first synchronises its stream, so issuing it on the compute
        stream would stall the host behind every queued kernel; the compute stream
        instead waits for the upload stream right before the next launch."""
        arr = np.ascontiguousarray(arr)
        tdt = self._TORCH_OF.get(arr.dtype)
        if dtype is not None or tdt is None or arr.nbytes > self._ARENA_MAX_ITEM:
            if not arr.flags.writeable:       # e.g. a read-only pandas view
                arr = arr.copy()
            t = torch.from_numpy(arr)
            if dtype is not None:
                t = t.to(dtype)
            main = self._main_stream()
            with torch.cuda.stream(self.h2d_stream):
                d = t.to(self.device, non_blocking=True)
            d.record_stream(main)
            self._h2d_dirty = True
            self.h2d_bytes += arr.nbytes
            return d
        if arr.nbytes == 0:
            return torch.empty(arr.shape, dtype=tdt, device=self.device)
        d = self._arena_take(arr.nbytes)
        _lib.check(self.lib.spx_upload_dev(C.c_void_p(d.data_ptr()), C.c_void_p(arr.ctypes.data),
                                           arr.nbytes, self._h2d_handle), 'upload')
        self._h2d_dirty = True
        self.h2d_bytes += arr.nbytes
        return d.view(tdt).view(arr.shape)

    def _dev_pack(self, arrays):
        """Upload several small arrays with ONE host->device copy; returns device
        views (16-byte aligned) in the same order."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        offs = []
        total = 0
        for a in arrays:
            total = (total + 15) & ~15
            offs.append(total)
            total += a.nbytes
        host = np.empty(max(total, 16), dtype=np.uint8)
        for a, o in zip(arrays, offs):
            host[o:o + a.nbytes] = a.view(np.uint8).ravel()
        d = self._dev(host)
        out = []
        for a, o in zip(arrays, offs):
            tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                   np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8}[a.dtype]
            out.append(d[o:o + a.nbytes].view(tdt).view(a.shape))
        return out

    def _mask_lists(self, ctx, rows_sel, offs, total, want):
        """Station index lists of mask rows on the device (spx_mask_lists_dev)."""
        out = torch.empty(max(int(total), 1), dtype=_I32, device=self.device)
        d_rows, d_off = self._dev_pack([np.asarray(rows_sel, dtype=np.int32),
                                        np.asarray(offs, dtype=np.int64)])
        _lib.check(self.lib.spx_mask_lists_dev(
            self._ptr(ctx['d_grp_mask']), ctx['n_stn'], self._ptr(d_rows), int(len(rows_sel)),
            self._ptr(d_off), int(want), self._ptr(out), self._stream()), 'mask_lists')
        self._count('launches')
        return out

    def _sync_uploads(self):
        """Make the compute stream wait for every upload issued so far."""
        if self._h2d_dirty:
            self._main_stream().wait_stream(self.h2d_stream)
            self._h2d_dirty = False

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _stream(self):
        self._sync_uploads()
        self._main_stream()
        return self._main_handle

    def _fetch_async(self, t):
        """Queue a device->pinned-host copy of a small tensor; the caller reads
        the returned host tensor after synchronising a later event."""
        t = t.contiguous()
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        nbytes = t.numel() * t.element_size()
        if nbytes:
            # SM copy into UVA-mapped pinned memory (not the DMA engine, which may be
            # busy for ~100 ms with a field download of the previous chunk)
            _lib.check(self.lib.spx_copy_to_mapped_host_dev(
                C.c_void_p(h.data_ptr()), C.c_void_p(t.data_ptr()), nbytes, self._stream()),
                'copy_to_mapped_host')
            self._count('launches')
        return h

    def _prof_begin(self):
        if not self.profile_gemm:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(self._main_stream())
        return e0

    def _prof_end(self, e0, name, bound, work):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(self._main_stream())
        self.kernel_events.append((name, bound, float(work), e0, e1))

    def _packed_downloader(self, t, decimals, depth=2):
        """The 2-byte transport (transfer.PackedDownloader) for a rounded f32 field that is
        large enough to matter, else None."""
        if (not self.packed_download or decimals is None or not (0 <= int(decimals) <= 9)
                or t.dtype != torch.float32 or t.dim() != 2 or t.numel() < (1 << 20)
                or t.shape[0] > 65535):
            return None
        dl = self._dl
        if (dl is None or dl.row_len != t.shape[1] or dl.max_rows < t.shape[0]
                or len(dl.slots) < depth or (self.transport and dl.codec != self.transport)):
            from .transfer import PackedDownloader
            dl = self._dl = PackedDownloader(self.device, t.shape[0], t.shape[1], depth=depth,
                                             codec=self.transport)
        return dl

    def round_and_stats(self, fld, decimals=None):
        """Round a device field [T, cells] in place (None: leave it) and return its
        per-step statistics as ndarray[5, T] (min, mean, max, std, count);
        spx_round_stats_dev."""
        n_rows, row_len = fld.shape
        out = np.empty((5, n_rows))
        dec = -1 if decimals is None else int(decimals)
        with torch.cuda.device(self.device):
            self._begin_call()
            for r0 in range(0, n_rows, 65535):
                r1 = min(n_rows, r0 + 65535)
                ws = torch.empty(max(1, self.lib.spx_round_stats_workspace(r1 - r0, row_len)),
                                 dtype=torch.uint8, device=self.device)
                sub = torch.empty((5, r1 - r0), dtype=_F64, device=self.device)
                _lib.check(self.lib.spx_round_stats_dev(
                    C.c_void_p(fld[r0:r1].data_ptr()), int(fld.dtype == _F64), r1 - r0, row_len,
                    fld.stride(0), dec, C.c_void_p(sub.data_ptr()), C.c_void_p(ws.data_ptr()),
                    self._stream()), 'round_stats')
                self._count('launches', 2)
                out[:, r0:r1] = sub.cpu().numpy()
        return out

    def trace_summary(self):
        """ms per entry point over self.trace (synchronises the device)."""
        torch.cuda.synchronize(self.device)
        out = {}
        for name, e0, e1 in self.trace:
            d = out.setdefault(name, {'ms': 0.0, 'n': 0})
            d['ms'] += e0.elapsed_time(e1)
            d['n'] += 1
        return out

    def _count(self, key, n=1):
        self.stats[key] = self.stats.get(key, 0) + n
        if key == 'launches':
            self.total_launches += n

    @contextlib.contextmanager
    def _phase(self, name):
        """Wall-clock per phase (with device syncs) when self.sync_timing is set;
        a debugging aid, off in normal runs."""
        if not self.sync_timing:
            yield
            return
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        yield
        torch.cuda.synchronize(self.device)
        self.timing[name] = self.timing.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)

    # ------------------------------------------------------------ public
    def interp_chunk(
            self, data, stn_xs, stn_ys, cell_xs, cell_ys, grid_shape, interp_args,
            vgs=None, cntn_idxs=None, drft_arrs=None, stns_drft=None,
            fld_beg_row=0, fld_end_row=None, neb_sel_mthd='all', n_nebs=None,
            min_var_thr=-np.inf, min_var_cut=None, max_var_cut=None,
            min_vg_val=0.0, est_var_flag=False, intrp_dtype=np.float32,
            return_device=False, n_pies=None):
        """Same contract as the reference's ``_get_all_interp_outputs`` reduced
        to arrays (see oracle/spinterp_oracle.py:interp_chunk for the argument
        meaning).  Returns ({label: ndarray[T, rows*cols]}, problem_steps)."""
        pend = self.submit_chunk(
            data, stn_xs, stn_ys, cell_xs, cell_ys, grid_shape, interp_args, vgs=vgs,
            cntn_idxs=cntn_idxs, drft_arrs=drft_arrs, stns_drft=stns_drft,
            fld_beg_row=fld_beg_row, fld_end_row=fld_end_row, neb_sel_mthd=neb_sel_mthd,
            n_nebs=n_nebs, min_var_thr=min_var_thr, min_var_cut=min_var_cut,
            max_var_cut=max_var_cut, min_vg_val=min_vg_val, est_var_flag=est_var_flag,
            intrp_dtype=intrp_dtype, n_pies=n_pies)
        return pend.result(to_host=not return_device)

    def submit_chunk(
            self, data, stn_xs, stn_ys, cell_xs, cell_ys, grid_shape, interp_args,
            vgs=None, cntn_idxs=None, drft_arrs=None, stns_drft=None,
            fld_beg_row=0, fld_end_row=None, neb_sel_mthd='all', n_nebs=None,
            min_var_thr=-np.inf, min_var_cut=None, max_var_cut=None,
            min_vg_val=0.0, est_var_flag=False, intrp_dtype=np.float32,
            round_decimals=None, field_stats=False, n_pies=None):
        """Queue every kernel of the chunk and return a PendingChunk.

        round_decimals / field_stats: the output stage of the reference on the device --
        fields rounded like ``np.round(flds, nmrl_prcn)`` before they leave the GPU and
        per-step min / mean / max / std / count in ``PendingChunk.field_stats``."""
        with torch.cuda.device(self.device):
            self._begin_call()
            self._arena_next_chunk()
            pend = self._interp_chunk(
                data, stn_xs, stn_ys, cell_xs, cell_ys, grid_shape, interp_args, vgs,
                cntn_idxs, drft_arrs, stns_drft, fld_beg_row, fld_end_row, neb_sel_mthd,
                n_nebs, min_var_thr, min_var_cut, max_var_cut, min_vg_val, est_var_flag,
                intrp_dtype, n_pies)
            pend.round_decimals = round_decimals
            pend.want_field_stats = bool(field_stats)
            return pend

    # ------------------------------------------------------------ impl
    def _interp_chunk(
            self, data, stn_xs, stn_ys, cell_xs, cell_ys, grid_shape, interp_args, vgs,
            cntn_idxs, drft_arrs, stns_drft, fld_beg_row, fld_end_row, neb_sel_mthd,
            n_nebs, min_var_thr, min_var_cut, max_var_cut, min_vg_val, est_var_flag,
            intrp_dtype, n_pies=None):
        self.stats = {}
        deferred = []
        self.timing = {}
        data = np.ascontiguousarray(data, dtype=np.float64)
        stn_xs = np.ascontiguousarray(stn_xs, dtype=np.float64)
        stn_ys = np.ascontiguousarray(stn_ys, dtype=np.float64)
        n_steps, n_stn = data.shape
        assert stn_xs.shape == stn_ys.shape == (n_stn,)
        if fld_end_row is None:
            fld_end_row = grid_shape[0]
        intrp_dtype = np.dtype(intrp_dtype)
        assert intrp_dtype in (np.dtype(np.float32), np.dtype(np.float64))
        out_f64 = int(intrp_dtype == np.dtype(np.float64))

        interp_types = [a[0] for a in interp_args]
        interp_labels = [a[2] for a in interp_args]
        krg_types = [t for t in ('OK', 'SK', 'EDK') if t in interp_types]
        edk_flag = 'EDK' in interp_types
        if krg_types:
            assert vgs is not None
            if not isinstance(vgs, list):
                vgs = list(vgs)
            vg_set = set(vgs)
            if any(type(v) is not str for v in vg_set):
                vgs = [str(v) for v in vgs]
                vg_set = set(vgs)
            assert len(vgs) == n_steps
            assert 'nan' not in vg_set, (            # steps.py:504-507
                'NaN VGs not allowed! Use Nugget or any other appropriate one!')
        if neb_sel_mthd not in ('all', 'nrst', 'pie'):
            raise NotImplementedError(f"neighbor selection '{neb_sel_mthd}'")
        # 'nrst' and 'pie' differ only in how a cell picks its n_nebs stations
        # (interp/grps.py:147-166 / :168-247); everything after that is shared
        nrst = neb_sel_mthd in ('nrst', 'pie')
        self._n_pies = 0
        if nrst:
            assert isinstance(n_nebs, (int, np.integer)) and n_nebs > 0
            if neb_sel_mthd == 'pie':
                assert isinstance(n_pies, (int, np.integer)) and 0 < n_pies <= n_nebs, (
                    'n_pies should be > 0 and <= n_neighbors!')
                self._n_pies = int(n_pies)
            if n_nebs >= n_stn:
                nrst = False            # interp/prepare.py:434-463 falls back to 'all'
            elif n_nebs > self.lib.spx_nrst_max_neighbors():
                raise NotImplementedError(
                    f'n_neighbors > {self.lib.spx_nrst_max_neighbors()} is not supported')
        ev_flag = bool(est_var_flag) and ('OK' in interp_types)
        if ev_flag:
            assert 'EST_VARS_OK' in interp_labels, 'est_var_flag needs an EST_VARS_OK label'

        # ---- cell subsetting, steps.py:512-568 ------------------------------
        # Depends on the grid only: cached per job (same coordinate arrays passed for
        # every chunk), together with the device copies and the bounding box.
        cell_xs = np.asarray(cell_xs)
        cell_ys = np.asarray(cell_ys)
        fld_n_cols = int(grid_shape[1])
        fld_beg_idx = fld_beg_row * fld_n_cols
        fld_end_idx = fld_end_row * fld_n_cols
        fld_size = (fld_end_row - fld_beg_row) * fld_n_cols
        # keyed by CONTENT (see _array_token), not by buffer addresses
        gkey = (self._array_token(cell_xs), self._array_token(cell_ys),
                None if cntn_idxs is None else self._array_token(cntn_idxs),
                fld_beg_row, fld_end_row, fld_n_cols)
        geo = self._geom_cache.get(gkey)
        if geo is None:
            if cntn_idxs is not None:
                whr = np.where(cntn_idxs)[0]
                sel = (whr >= fld_beg_idx) & (whr < fld_end_idx)
                msh_idxs = np.arange(whr.size)[sel]
                out_pos = (whr[sel] - fld_beg_idx).astype(np.int32)
                dst_xs = np.ascontiguousarray(cell_xs[msh_idxs], dtype=np.float64)
                dst_ys = np.ascontiguousarray(cell_ys[msh_idxs], dtype=np.float64)
            else:
                msh_idxs = None
                out_pos = None
                dst_xs = np.ascontiguousarray(cell_xs[fld_beg_idx:fld_end_idx], dtype=np.float64)
                dst_ys = np.ascontiguousarray(cell_ys[fld_beg_idx:fld_end_idx], dtype=np.float64)
            empty = dst_xs.shape[0] == 0
            geo = dict(msh_idxs=msh_idxs, out_pos=out_pos, dst_xs=dst_xs, dst_ys=dst_ys,
                       d_cell_x=None if empty else self._dev_keep(dst_xs),
                       d_cell_y=None if empty else self._dev_keep(dst_ys),
                       d_pos=(self._dev_keep(out_pos) if (out_pos is not None and not empty)
                              else None),
                       bbox=(0.0, 0.0, 0.0, 0.0) if empty else (
                           float(dst_xs.min()), float(dst_xs.max()), float(dst_ys.min()),
                           float(dst_ys.max())),
                       fp=gkey)
            while len(self._geom_cache) >= 4:
                self._geom_cache.pop(next(iter(self._geom_cache)))
            self._geom_cache[gkey] = geo
        out_pos, dst_xs, dst_ys = geo['out_pos'], geo['dst_xs'], geo['dst_ys']
        if dst_xs.shape[0] == 0:
            # no cell of this grid-row chunk is selected: the fields stay NaN
            # (steps.py:659-663 pre-fills them) and nothing is computed
            tdt0 = torch.float64 if out_f64 else torch.float32
            flds0 = {lab: torch.full((n_steps, fld_size), float('nan'), dtype=tdt0,
                                     device=self.device) for lab in interp_labels}
            n_av0 = np.isfinite(data).sum(axis=1)
            return PendingChunk(self, flds0, [int(s_) for s_ in np.where(n_av0 == 0)[0]], [],
                                dict(self.stats))
        if drft_arrs is not None:
            drft_arrs = (drft_arrs[:, geo['msh_idxs']] if geo['msh_idxs'] is not None
                         else drft_arrs[:, fld_beg_idx:fld_end_idx])
        n_cells = int(dst_xs.shape[0])

        if nrst:
            # stations never among the n_nebs nearest of any cell are dropped up-front
            # (interp/steps.py:592-608, quirk Q9)
            # depends on the coordinates only: once per job and geometry, not per chunk
            pkey = (stn_xs.tobytes(), stn_ys.tobytes(), int(n_nebs), int(self._n_pies))
            prune = geo.setdefault('nrst_prune', {})
            tke = prune.get(pkey)
            if tke is None:
                nb0, _ = self._topk(self._dev_const(stn_xs), self._dev_const(stn_ys), n_stn, None,
                                    geo['d_cell_x'], geo['d_cell_y'], n_cells, int(n_nebs))
                tke = torch.unique(nb0).cpu().numpy()
                while len(prune) >= 4:
                    prune.pop(next(iter(prune)))
                prune[pkey] = tke
            if tke.size != n_stn:
                stn_xs = np.ascontiguousarray(stn_xs[tke])
                stn_ys = np.ascontiguousarray(stn_ys[tke])
                data = np.ascontiguousarray(data[:, tke])
                if stns_drft is not None:
                    stns_drft = np.ascontiguousarray(np.asarray(stns_drft)[tke])
                n_stn = int(tke.size)

        d_stn_x = self._dev_const(stn_xs)
        d_stn_y = self._dev_const(stn_ys)
        d_cell_x, d_cell_y, d_pos = geo['d_cell_x'], geo['d_cell_y'], geo['d_pos']
        ctx = _LazyCtx(
            n_steps=n_steps, n_stn=n_stn, n_cells=n_cells, fld_size=fld_size, out_f64=out_f64,
            d_stn_x=d_stn_x, d_stn_y=d_stn_y, d_cell_x=d_cell_x, d_cell_y=d_cell_y, d_pos=d_pos,
            out_pos=out_pos, dst_xs=dst_xs, dst_ys=dst_ys, stn_xs=stn_xs, stn_ys=stn_ys,
            has_lo=int(min_var_cut is not None), has_hi=int(max_var_cut is not None),
            lo=float(min_var_cut) if min_var_cut is not None else 0.0,
            hi=float(max_var_cut) if max_var_cut is not None else 0.0,
            min_vg_val=float(min_vg_val), nnb_cache={}, bbox=geo['bbox'], geom_fp=geo['fp'],
            geom_key=gkey)
        tdtype = torch.float64 if out_f64 else torch.float32

        # ---- one native call for the whole kriging label (csrc/spx_chunk.cu) ---------
        fast = None
        if (self.native_submit and self.native_plan and self.downdate and self.ginv_cache
                and not nrst and not ev_flag and krg_types == ['OK'] and len(vg_set) == 1
                and not self.sync_timing and not self.trace_launches):
            fast = self._fast_submit(ctx, data, vgs[0], float(min_var_thr), tdtype,
                                     interp_labels[interp_types.index('OK')])

        t_host0 = time.perf_counter()
        if fast is not None:
            grp_of_step, grp_bits, grp_n, grp_first, n_avail, steps_flags = fast['groups']
            data_upload = None
        else:
            # The data block goes to the device from a helper thread: a host -> device copy
            # from pageable memory blocks its caller for the whole staging copy (0.4 ms for
            # 5 MB), and both that call and the native planner below run without the GIL.
            d_data, data_upload = self._dev_threaded(data)
            ctx['d_data'] = d_data
            # availability groups (grps.py:57-101) and per-step flags (steps.py:760-765) in
            # one pass over the data (native: csrc/spx_plan.cu); byte masks are expanded
            # only if a path asks for them
            grp_of_step, grp_bits, grp_n, grp_first, n_avail, steps_flags = _lib.avail_groups(
                data, float(min_var_thr), want_mask=False)
        n_grps = int(grp_bits.shape[0])
        problem_steps = [int(s) for s in np.where(n_avail == 0)[0]]   # steps.py:677-688

        def ref_means_of(idx):                                       # steps.py:276, on demand
            return np.nansum(data[idx], axis=1) / n_avail[idx]

        single_val_of = lambda idx: np.nansum(data[idx], axis=1)    # noqa: E731  n_avail == 1

        self.timing['host_groups'] = 1e3 * (time.perf_counter() - t_host0)
        if data_upload is not None:
            data_upload.result()               # re-raises a failure of the copy
        ctx.update(grp_of_step=grp_of_step, grp_n=grp_n, n_avail=n_avail, n_grps=n_grps,
                   grp_first=grp_first)

        def _d_data0():
            d = ctx['d_data']
            self._sync_uploads()
            return torch.nan_to_num(d, nan=0.0, posinf=float('inf'), neginf=float('-inf'))

        def _d_grp_mask():
            # availability masks on the device (+ one all-ones row = "every station")
            return self._dev(np.concatenate(
                [ctx['grp_mask'].view(np.uint8), np.ones((1, n_stn), dtype=np.uint8)], axis=0))

        ctx.lazy = dict(grp_mask=lambda: _lib.unpack_group_bits(grp_bits, n_stn),
                        d_data0=_d_data0, d_grp_mask=_d_grp_mask,
                        d_data=lambda: self._dev(data))

        # NaN marks cells outside the mask and steps without stations
        # (steps.py:659-663); when every cell of every step is written the 4-byte
        # per cell-step prefill is skipped
        full_cover = (out_pos is None) and bool((n_avail >= 1).all())
        flds = {}
        for lab in interp_labels:
            if fast is not None and lab == fast['label']:
                flds[lab] = fast['out']
                if out_pos is None and not full_cover:
                    # allocated before the availability was known: NaN rows written now
                    none_steps = np.where(n_avail == 0)[0]
                    self._fill_rows(ctx, fast['out'], none_steps,
                                    np.full(none_steps.size, np.nan), clamp=False)
            elif full_cover and lab != 'EST_VARS_OK':
                flds[lab] = torch.empty((n_steps, fld_size), dtype=tdtype, device=self.device)
            else:
                flds[lab] = torch.full((n_steps, fld_size), float('nan'), dtype=tdtype,
                                       device=self.device)

        # steps that bypass interpolation for every method
        single_steps = np.where(n_avail == 1)[0]                    # steps.py:282-283

        for i, itype in enumerate(interp_types):
            lab = interp_labels[i]
            if lab == 'EST_VARS_OK':
                continue
            out = flds[lab]
            if single_steps.size:
                self._fill_rows(ctx, out, single_steps, single_val_of(single_steps))
            multi = n_avail >= 2
            if fast is not None and lab == fast['label']:
                # kriged steps are already queued (one variogram, not nugget-only)
                if fast['res'].n_mean:
                    mean_steps = np.where(multi & ~steps_flags)[0]   # steps.py:325-331
                    self._fill_rows(ctx, out, mean_steps, ref_means_of(mean_steps))
                deferred.append(self._fast_deferred(ctx, fast, out, multi & steps_flags,
                                                    problem_steps))
            elif itype == 'NNB':
                self._nnb_label(ctx, out, np.where(multi)[0])
            elif itype == 'IDW' and nrst:
                self._nrst(ctx, out, 'IDW', np.where(multi)[0], int(n_nebs),
                           idw_exp=float(interp_args[i][3]), min_var_thr=float(min_var_thr))
            elif itype == 'IDW':
                mean_steps = np.where(multi & ~steps_flags)[0]       # steps.py:312-313
                if mean_steps.size:
                    self._fill_rows(ctx, out, mean_steps, ref_means_of(mean_steps))
                self._idw(ctx, out, np.where(multi & steps_flags)[0], float(interp_args[i][3]))
            elif itype in ('OK', 'SK', 'EDK') and nrst:
                uniq_vgs = list(dict.fromkeys(vgs))
                vg_id = {v: k for k, v in enumerate(uniq_vgs)}
                nug = np.array([check_full_nuggetness(v, min_vg_val) for v in uniq_vgs])
                step_vg = np.fromiter(map(vg_id.__getitem__, vgs), dtype=np.int32, count=len(vgs))
                self._nrst(ctx, out, itype, np.where(multi)[0], int(n_nebs), step_vg=step_vg,
                           uniq_vgs=uniq_vgs, nug=nug, min_var_thr=float(min_var_thr),
                           drft_arrs=drft_arrs if itype == 'EDK' else None,
                           stns_drft=stns_drft if itype == 'EDK' else None,
                           ev_out=flds['EST_VARS_OK'] if (ev_flag and itype == 'OK') else None)
            elif itype in ('OK', 'SK', 'EDK'):
                uniq_vgs = list(dict.fromkeys(vgs))
                vg_id = {v: k for k, v in enumerate(uniq_vgs)}
                nug = np.array([check_full_nuggetness(v, min_vg_val) for v in uniq_vgs])
                step_vg = np.fromiter(map(vg_id.__getitem__, vgs), dtype=np.int32, count=len(vgs))
                bypass = (~steps_flags) | nug[step_vg]               # steps.py:325-331
                mean_steps = np.where(multi & bypass)[0]
                if mean_steps.size:
                    self._fill_rows(ctx, out, mean_steps, ref_means_of(mean_steps))
                ev_out = flds['EST_VARS_OK'] if (ev_flag and itype == 'OK') else None
                if ev_out is not None and mean_steps.size:          # steps.py:329
                    self._fill_rows(ctx, ev_out, mean_steps, np.zeros(mean_steps.size),
                                    clamp=False)
                fn = self._krige(ctx, out, itype, np.where(multi & ~bypass)[0], step_vg,
                                 uniq_vgs, drft_arrs if itype == 'EDK' else None,
                                 stns_drft if itype == 'EDK' else None, problem_steps,
                                 ev_out=ev_out)
                if fn is not None:
                    deferred.append(fn)
            else:
                raise NotImplementedError(itype)

        return PendingChunk(self, flds, problem_steps, deferred, dict(self.stats))

    # ------------------------------------------------------------ pieces
    def _fill_rows(self, ctx, out, steps, vals, clamp=True):
        d_vals = self._dev(np.asarray(vals, dtype=np.float64))
        d_rows = self._dev(np.asarray(steps, dtype=np.int32))
        _lib.check(self.lib.spx_fill_rows_dev(
            self._ptr(d_vals), self._ptr(d_rows), len(steps), ctx['n_cells'], self._ptr(ctx['d_pos']),
            self._ptr(out), ctx['fld_size'], ctx['out_f64'], ctx['has_lo'] if clamp else 0,
            ctx['has_hi'] if clamp else 0, ctx['lo'], ctx['hi'], self._stream()), 'fill_rows')
        self._count('launches')

    def _nnb_index(self, ctx, grp_ids, cells=None):
        """nnb[len(grp_ids), n_cells] int32 for the listed groups (cached for
        the full cell set)."""
        key = (tuple(int(g) for g in grp_ids), None if cells is None else cells.tobytes())
        if key in ctx['nnb_cache']:
            return ctx['nnb_cache'][key]
        mask = self._dev(ctx['grp_mask'][np.asarray(grp_ids)].astype(np.uint8))
        if cells is None:
            cx, cy, nc = ctx['d_cell_x'], ctx['d_cell_y'], ctx['n_cells']
        else:
            cx = self._dev(ctx['dst_xs'][cells])
            cy = self._dev(ctx['dst_ys'][cells])
            nc = int(cells.size)
        nnb = torch.empty((len(grp_ids), nc), dtype=_I32, device=self.device)
        if len(grp_ids) >= 4 and ctx['n_stn'] > self.lib.spx_nnb_candidates_width():
            # candidate lists once per (chunk, cell set), then a cheap pass per group
            ckey = ('cand', None if cells is None else cells.tobytes())
            cand = ctx['nnb_cache'].get(ckey)
            if cand is None:
                cand = torch.empty((nc, self.lib.spx_nnb_candidates_width()), dtype=_I32,
                                   device=self.device)
                _lib.check(self.lib.spx_nnb_candidates_dev(
                    self._ptr(ctx['d_stn_x']), self._ptr(ctx['d_stn_y']), ctx['n_stn'],
                    self._ptr(cx), self._ptr(cy), nc, self._ptr(cand), self._stream()),
                    'nnb_candidates')
                self._count('launches')
                ctx['nnb_cache'][ckey] = cand
            _lib.check(self.lib.spx_nnb_index_cand_dev(
                self._ptr(ctx['d_stn_x']), self._ptr(ctx['d_stn_y']), ctx['n_stn'],
                self._ptr(mask), len(grp_ids), self._ptr(cx), self._ptr(cy), nc,
                self._ptr(cand), self._ptr(nnb), self._stream()), 'nnb_index_cand')
        else:
            _lib.check(self.lib.spx_nnb_index_dev(
                self._ptr(ctx['d_stn_x']), self._ptr(ctx['d_stn_y']), ctx['n_stn'],
                self._ptr(mask), len(grp_ids), self._ptr(cx), self._ptr(cy), nc,
                self._ptr(nnb), self._stream()), 'nnb_index')
        self._count('launches')
        ctx['nnb_cache'][key] = nnb
        return nnb

    def _nnb_gather(self, ctx, out, nnb, row_step, row_grp_slot, fail=None, row_fail=None,
                    n_cells=None, d_pos=None):
        n_cells = ctx['n_cells'] if n_cells is None else n_cells
        d_pos = ctx['d_pos'] if d_pos is None else d_pos
        d_step = self._dev(np.asarray(row_step, dtype=np.int32))
        d_grp = self._dev(np.asarray(row_grp_slot, dtype=np.int32))
        d_rf = self._dev(np.asarray(row_fail, dtype=np.int32)) if row_fail is not None else None
        _lib.check(self.lib.spx_nnb_gather_dev(
            self._ptr(ctx['d_data']), ctx['n_stn'], self._ptr(nnb), self._ptr(d_step),
            self._ptr(d_grp), self._ptr(d_step), len(row_step), self._ptr(fail), self._ptr(d_rf),
            n_cells, self._ptr(d_pos), self._ptr(out), ctx['fld_size'], ctx['out_f64'],
            ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi'], self._stream()), 'nnb_gather')
        self._count('launches')

    def _nnb_label(self, ctx, out, steps):
        """interp/steps.py:285-291."""
        if not steps.size:
            return
        grp_of_step = ctx['grp_of_step']
        grps = np.unique(grp_of_step[steps])
        # bound the index buffer: groups in batches
        per_grp = ctx['n_cells'] * 4
        batch = max(1, int(self.aux_limit // max(per_grp, 1)))
        for b0 in range(0, grps.size, batch):
            gb = grps[b0:b0 + batch]
            slot_of = np.full(ctx['n_grps'], -1, dtype=np.int32)
            slot_of[gb] = np.arange(gb.size, dtype=np.int32)
            st = steps[slot_of[grp_of_step[steps]] >= 0]
            nnb = self._nnb_index(ctx, gb)
            self._nnb_gather(ctx, out, nnb, st, slot_of[grp_of_step[st]])

    def _gemm(self, ctx, *, coef, n_rows, kpad, n_border, gen, epi, row_dst, row_aux=None,
              out=None, aux=None, vg=None, covar_flag=0, idw_exp=0.0, dist_scale=1.0,
              cell_drift=None, quad_slot=0):
        g = _lib.spx_gemm()
        g.coef = coef.data_ptr()
        g.n_rows = int(n_rows)
        g.kpad = int(kpad)
        g.n_stn = ctx['n_stn']
        g.n_border = int(n_border)
        g.stn_x = ctx['d_stn_x'].data_ptr()
        g.stn_y = ctx['d_stn_y'].data_ptr()
        g.cell_x = ctx['d_cell_x'].data_ptr()
        g.cell_y = ctx['d_cell_y'].data_ptr()
        g.n_cells = ctx['n_cells']
        g.cell_drift = cell_drift.data_ptr() if cell_drift is not None else None
        g.gen = gen
        g.covar_flag = int(covar_flag)
        if vg is not None:
            g.vg = vg
        g.min_vg_val = ctx['min_vg_val']
        g.idw_exp = float(idw_exp)
        g.dist_scale = float(dist_scale)
        g.epi = epi
        g.row_dst = row_dst.data_ptr()
        g.row_aux = row_aux.data_ptr() if row_aux is not None else None
        g.out = out.data_ptr() if out is not None else None
        g.out_ld = ctx['fld_size']
        g.out_f64 = ctx['out_f64']
        g.cell_pos = ctx['d_pos'].data_ptr() if ctx['d_pos'] is not None else None
        g.aux = aux.data_ptr() if aux is not None else None
        g.has_lo, g.has_hi, g.lo, g.hi = ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi']
        g.quad_slot = int(quad_slot)
        if self.profile_gemm:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(self.lib.spx_estimate_gemm_dev(C.byref(g), self._stream()), 'estimate_gemm')
        if self.profile_gemm:
            e1.record()
            self.gemm_events.append((e0, e1))
            self.kernel_events.append(('k_estimate_gemm', 'tensor',
                                       2.0 * int(n_rows) * int(kpad) * ctx['n_cells'], e0, e1))
        self._count('launches')
        self._count('gemm_launches')
        self._count('gemm_flop', 2 * int(n_rows) * int(kpad) * ctx['n_cells'])

    # ---- local estimator (compact support) ---------------------------------
    def _local_plan(self, ctx, vg_strs):
        """[(R, F_var, F_cov)] per variogram if every one is compactly supported
        (only Nug / Sph / Lin terms) and sparse enough for the local estimator, else
        None."""
        x0, x1, y0, y1 = ctx['bbox']
        area = max((x1 - x0) * (y1 - y0), 1e-300)
        plan = []
        for vg_s in vg_strs:
            terms = _lib.parse_vg_str(vg_s)
            R = 0.0
            tot = 0.0
            for t, sill, rng in terms:
                if t not in (1, 2, 4):
                    return None
                tot = tot + sill                  # same order as pyx:192-201 accumulates
                if t != 1:
                    R = max(R, rng)
            if R <= 0.0:
                return None
            if ctx['n_stn'] * math.pi * R * R / area > self.local_max_near:
                return None
            plan.append((R, tot))
        return plan

    def _local_neighbours(self, ctx, K, vg_s, plan_entry, transient=None):
        """Stations within the range of every cell and their (vg - F) values;
        depends on the geometry and the variogram only -> cached across chunks.

        transient (a dict shared by the variograms of ONE call, see _estimate): a chunk
        with more variograms than the cache holds (per-step variogram series) builds the
        tables of each variogram into the same buffers, one after the other on the stream,
        with uploads from the chunk arena, a capacity fixed once from the largest range
        (stations within R are a subset of those within R_max) and no host synchronisation
        per variogram."""
        R, tot = plan_entry
        covar = int(K.kind == 1)
        mv = ctx['min_vg_val']
        F = 0.0 if covar else tot
        if F <= mv:
            F = 0.0                               # pyx:203-216
        fp = (ctx['geom_fp'], ctx['bbox'], ctx['n_stn'], ctx['stn_xs'].tobytes(),
              ctx['stn_ys'].tobytes())
        key = (vg_s, covar, mv, fp)
        hit = self._local_cache.get(key)
        if hit is not None:
            self._count('local_cache_hits')
            return hit
        sx, sy = ctx['stn_xs'], ctx['stn_ys']
        x0, y0 = sx.min() - R, sy.min() - R
        nbx = int(math.floor((sx.max() + R - x0) / R)) + 1
        nby = int(math.floor((sy.max() + R - y0) / R)) + 1
        b = (np.floor((sy - y0) / R).astype(np.int64) * nbx
             + np.floor((sx - x0) / R).astype(np.int64))
        order = np.argsort(b, kind='stable').astype(np.int32)
        bin_start = np.searchsorted(b[order], np.arange(nbx * nby + 1)).astype(np.int32)
        n_cells = ctx['n_cells']
        if transient is not None and transient.get('cap'):
            d_bs, d_bo = self._dev(bin_start), self._dev(order)     # chunk arena
        else:
            d_bs, d_bo = self._dev_keep(bin_start), self._dev_keep(order)   # cached with the tables
        tiles = ctx['n_stn'] <= 65536 and self.local_tiles
        n_tiles = (n_cells + _lib.SPX_LOCAL_TILE - 1) // _lib.SPX_LOCAL_TILE

        def buffers(cap):
            out = dict(cnt=torch.empty(n_cells, dtype=_I32, device=self.device),
                       idx=torch.empty((cap, n_cells), dtype=_I32, device=self.device),
                       val=torch.empty((cap, n_cells), dtype=_F64, device=self.device))
            if tiles:
                out.update(tile_cnt=torch.empty(n_tiles, dtype=_I32, device=self.device),
                           tile_stn=torch.empty((n_tiles, _lib.SPX_LOCAL_TILE_CAP), dtype=_I32,
                                                device=self.device),
                           slot=torch.empty((cap, n_cells), dtype=torch.uint8, device=self.device))
            return out

        def struct_for(buf, cap):
            L = _lib.spx_local()
            L.stn_x, L.stn_y = ctx['d_stn_x'].data_ptr(), ctx['d_stn_y'].data_ptr()
            L.bin_start, L.bin_stn = d_bs.data_ptr(), d_bo.data_ptr()
            L.x0, L.y0, L.inv_bin, L.nbx, L.nby = float(x0), float(y0), 1.0 / R, nbx, nby
            L.R, L.F = float(R), float(F)
            L.cell_x, L.cell_y = ctx['d_cell_x'].data_ptr(), ctx['d_cell_y'].data_ptr()
            L.n_cells = n_cells
            L.cap = cap
            L.cnt, L.idx, L.val = buf['cnt'].data_ptr(), buf['idx'].data_ptr(), buf['val'].data_ptr()
            L.vg = _lib.make_vg(vg_s)
            L.covar_flag = covar
            L.min_vg_val = mv
            return L

        if transient is not None and transient.get('cap'):
            cap, buf = transient['cap'], transient['buf']
            mx = transient['max_near']
            L = struct_for(buf, cap)
            _lib.check(self.lib.spx_local_build_dev(C.byref(L), self._stream()), 'local_build')
            self._count('launches')
        else:
            cap = 8
            while True:
                buf = buffers(cap)
                L = struct_for(buf, cap)
                _lib.check(self.lib.spx_local_build_dev(C.byref(L), self._stream()), 'local_build')
                self._count('launches')
                mx = int(buf['cnt'].max().item())
                if mx <= cap:
                    break
                cap = mx
        # every tensor whose address sits in the cached struct stays referenced with it
        keep = [buf, d_bs, d_bo, ctx['d_stn_x'], ctx['d_stn_y'], ctx['d_cell_x'], ctx['d_cell_y']]
        if tiles:
            # distinct near stations per tile of 256 cells: the streamlined kernel stages
            # their coefficient slices in shared memory
            L.n_stn = ctx['n_stn']
            L.tile_cnt, L.tile_stn, L.slot = (buf['tile_cnt'].data_ptr(), buf['tile_stn'].data_ptr(),
                                              buf['slot'].data_ptr())
            _lib.check(self.lib.spx_local_tiles_dev(C.byref(L), self._stream()), 'local_tiles')
            self._count('launches')
        res = dict(struct=L, F=F, cap=cap, keep=tuple(keep), max_near=mx)
        self.stats['local_max_near'] = mx
        if transient is not None:
            if not transient.get('cap'):         # the R_max build of the call: sizes the rest
                transient.update(cap=cap, buf=buf, max_near=mx)
            return res
        while len(self._local_cache) >= 4:
            self._local_cache.pop(next(iter(self._local_cache)))
        self._local_cache[key] = res
        return res

    def _local_aux(self, ctx, K, vg_s, plan_entry, coef, n_rows, d_slots, aux):
        """aux[d_slots[r], cell] = sum_k coef[r, k] * vg(dist(k, cell)) for ROW-MAJOR coef rows
        through the local estimator (f64, no clamp): the per-cell sum(lambda) of
        steps.py:418 for systems whose variogram is compactly supported."""
        nbr = self._local_neighbours(ctx, K, vg_s, plan_entry)
        coef2d = coef.view(-1, K.kpad)
        base = nbr['F'] * coef2d[:n_rows, :ctx['n_stn']].sum(dim=1)
        if K.n_border >= 1:
            base = base + coef2d[:n_rows, ctx['n_stn']]
        base = base.contiguous()
        L = nbr['struct']
        L.coef = coef2d.data_ptr()
        L.base = base.data_ptr()
        L.n_rows = int(n_rows)
        L.kpad, L.n_stn, L.n_drifts = K.kpad, ctx['n_stn'], 0
        L.cell_drift = None
        L.row_dst = d_slots.data_ptr()
        L.out = aux.data_ptr()
        L.out_ld = ctx['n_cells']
        L.out_f64 = 1
        L.cell_pos = None
        L.has_lo = L.has_hi = 0
        L.lo = L.hi = 0.0
        L.rows_all_valid = 1
        L.coef_t, L.coef_t_ld = None, 0
        _lib.check(self.lib.spx_estimate_local_dev(C.byref(L), self._stream()), 'estimate_local')
        self._count('launches')
        self._count('local_aux_rows', int(n_rows))

    # ---- 'nrst' neighbour selection ---------------------------------------
    def _topk(self, d_sx, d_sy, n_stn, d_mask, d_cx, d_cy, n_cells, k):
        nb = torch.empty((n_cells, k), dtype=_I32, device=self.device)
        hsh = torch.empty(n_cells, dtype=_I64, device=self.device)
        if self._n_pies:
            _lib.check(self.lib.spx_pie_select_dev(
                self._ptr(d_sx), self._ptr(d_sy), n_stn, self._ptr(d_mask), self._ptr(d_cx),
                self._ptr(d_cy), n_cells, k, self._n_pies, self._ptr(nb), self._ptr(hsh),
                self._stream()), 'pie_select')
        else:
            _lib.check(self.lib.spx_nrst_topk_dev(
                self._ptr(d_sx), self._ptr(d_sy), n_stn, self._ptr(d_mask), self._ptr(d_cx),
                self._ptr(d_cy), n_cells, k, self._ptr(nb), self._ptr(hsh), self._stream()),
                'nrst_topk')
        self._count('launches')
        return nb, hsh

    def neighbor_indices(self, stn_xs, stn_ys, cell_xs, cell_ys, neb_sel_mthd, n_nebs,
                         n_pies=None, avail=None):
        """``get_neb_idxs_and_grps`` of the reference (interp/grps.py:249-288) on the GPU:
        the neighbour row of every cell (ascending station indices; 'nrst': the n_nebs
        nearest, 'pie': round-robin over n_pies angular sectors) and the groups of cells
        with identical rows -- groups in first-occurrence order, members ascending.
        avail: optional bool [n_stn], only these stations are candidates.
        Returns (all_neb_idxs int64 [n_cells, k], [int64 arrays])."""
        assert neb_sel_mthd in ('nrst', 'pie')
        stn_xs = np.ascontiguousarray(stn_xs, dtype=np.float64)
        stn_ys = np.ascontiguousarray(stn_ys, dtype=np.float64)
        cell_xs = np.ascontiguousarray(cell_xs, dtype=np.float64)
        cell_ys = np.ascontiguousarray(cell_ys, dtype=np.float64)
        n_stn, n_cells = int(stn_xs.size), int(cell_xs.size)
        n_av = n_stn if avail is None else int(np.count_nonzero(avail))
        k = int(min(n_nebs, n_av))
        with torch.cuda.device(self.device):
            self._begin_call()
            self._n_pies = int(n_pies) if neb_sel_mthd == 'pie' else 0
            d_mask = None if avail is None else self._dev(np.asarray(avail).astype(np.uint8))
            nb, hsh = self._topk(self._dev(stn_xs), self._dev(stn_ys), n_stn, d_mask,
                                 self._dev(cell_xs), self._dev(cell_ys), n_cells, k)
            uh, inv = torch.unique(hsh, return_inverse=True)
            rep = torch.full((int(uh.numel()),), n_cells, dtype=_I64, device=self.device)
            rep.scatter_reduce_(0, inv, torch.arange(n_cells, device=self.device), 'amin')
            order = torch.argsort(rep)                      # first-occurrence order
            rank = torch.empty_like(order)
            rank[order] = torch.arange(order.numel(), device=self.device)
            grp_of_cell = rank[inv].cpu().numpy()
            idxs = nb.cpu().numpy().astype(np.int64)
        by = np.argsort(grp_of_cell, kind='stable')
        bounds = np.searchsorted(grp_of_cell[by], np.arange(int(uh.numel()) + 1))
        grps = [by[bounds[g]:bounds[g + 1]].astype(np.int64) for g in range(int(uh.numel()))]
        return idxs, grps

    def _nrst_groups(self, ctx, g, n_nebs):
        """Neighbour rows and cell groups of availability group g, cached per
        chunk (interp/grps.py:249-288 is called per group, steps.py:732)."""
        key = ('nrst', int(g))
        if key in ctx['nnb_cache']:
            return ctx['nnb_cache'][key]
        k = int(min(n_nebs, ctx['grp_n'][g]))
        d_mask = self._dev(ctx['grp_mask'][g].astype(np.uint8))
        nb, hsh = self._topk(ctx['d_stn_x'], ctx['d_stn_y'], ctx['n_stn'], d_mask,
                             ctx['d_cell_x'], ctx['d_cell_y'], ctx['n_cells'], k)
        uh, inv = torch.unique(hsh, return_inverse=True)
        n_grp = int(uh.numel())
        rep = torch.full((n_grp,), ctx['n_cells'], dtype=_I64, device=self.device)
        rep.scatter_reduce_(0, inv, torch.arange(ctx['n_cells'], device=self.device), 'amin')
        nbu = nb.index_select(0, rep).contiguous()
        res = (k, nb, nbu, inv.to(_I32).contiguous(), n_grp)
        ctx['nnb_cache'] = {kk: vv for kk, vv in ctx['nnb_cache'].items()
                            if not (isinstance(kk, tuple) and kk and kk[0] == 'nrst')}
        ctx['nnb_cache'][key] = res
        return res

    def _nrst(self, ctx, out, itype, steps, n_nebs, step_vg=None, uniq_vgs=None, nug=None,
              idw_exp=0.0, min_var_thr=-np.inf, drft_arrs=None, stns_drft=None, ev_out=None):
        """Kriging / IDW with the k nearest available stations per cell
        (interp/steps.py:740-833 with neb_sel_mthd == 'nrst')."""
        if not steps.size:
            return
        lib = self.lib
        n_cells, n_stn = ctx['n_cells'], ctx['n_stn']
        grp_of_step = ctx['grp_of_step']
        kind = _lib.KRG_KINDS.get(itype, 0)
        n_drifts = 0
        d_cell_drift = d_stn_drift = None
        if itype == 'EDK':
            n_drifts = int(np.asarray(stns_drft).shape[1])
            d_cell_drift = self._dev(np.ascontiguousarray(drft_arrs, dtype=np.float64))
            d_stn_drift = self._dev(np.ascontiguousarray(stns_drft, dtype=np.float64))
        n_border = {'OK': 1, 'SK': 0, 'EDK': 1 + n_drifts}.get(itype, 0)

        for g in np.unique(grp_of_step[steps]):
            st_g = steps[grp_of_step[steps] == g]
            k, nb, nbu, cell_grp, n_grp = self._nrst_groups(ctx, g, n_nebs)
            N = _lib.spx_nrst()
            N.n_grp = n_grp
            N.n_cells = n_cells
            N.k, N.n_border, N.n_drifts, N.kind, N.n_stn = k, n_border, n_drifts, kind, n_stn
            N.nbu = nbu.data_ptr()
            N.cell_grp = cell_grp.data_ptr()
            N.stn_x = ctx['d_stn_x'].data_ptr()
            N.stn_y = ctx['d_stn_y'].data_ptr()
            N.stn_drift = d_stn_drift.data_ptr() if d_stn_drift is not None else None
            N.cell_x = ctx['d_cell_x'].data_ptr()
            N.cell_y = ctx['d_cell_y'].data_ptr()
            N.cell_drift = d_cell_drift.data_ptr() if d_cell_drift is not None else None
            N.min_vg_val = ctx['min_vg_val']
            N.data = ctx['d_data'].data_ptr()
            N.min_var_thr = float(min_var_thr)
            N.cell_pos = ctx['d_pos'].data_ptr() if ctx['d_pos'] is not None else None
            N.out = out.data_ptr()
            N.out_ld = ctx['fld_size']
            N.out_f64 = ctx['out_f64']
            N.has_lo, N.has_hi, N.lo, N.hi = ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi']
            N.idw_exp = float(idw_exp)
            if itype == 'IDW':
                d_steps = self._dev(st_g.astype(np.int32))
                N.steps = d_steps.data_ptr()
                N.n_t = int(st_g.size)
                _lib.check(lib.spx_nrst_idw_dev(C.byref(N), self._ptr(nb), self._stream()),
                           'nrst_idw')
                self._count('launches')
                continue
            m = k + n_border
            t_max = max(1, int(self.aux_limit // (n_grp * m * 8)) - 1)
            for v in np.unique(step_vg[st_g]):
                st_v = st_g[step_vg[st_g] == v]
                N.vg = _lib.make_vg(uniq_vgs[int(v)])
                for b0 in range(0, st_v.size, t_max):
                    sb = st_v[b0:b0 + t_max]
                    n_t = int(sb.size)
                    d_steps = self._dev(sb.astype(np.int32))
                    d_byp = self._dev(np.full(n_t, int(bool(nug[int(v)])), dtype=np.uint8))
                    coef = torch.empty((n_grp, n_t + 1, m), dtype=_F64, device=self.device)
                    ovr = torch.empty((n_grp, n_t), dtype=_F64, device=self.device)
                    info = torch.zeros(n_grp, dtype=_I32, device=self.device)
                    N.steps = d_steps.data_ptr()
                    N.n_t = n_t
                    N.step_bypass = d_byp.data_ptr()
                    N.coef, N.ovr, N.info = coef.data_ptr(), ovr.data_ptr(), info.data_ptr()
                    if ev_out is None:
                        N.inv, N.ev_out, N.u_beg, N.u_end = None, None, 0, 0
                        _lib.check(lib.spx_nrst_solve_dev(C.byref(N), self._stream()), 'nrst_solve')
                        _lib.check(lib.spx_nrst_krige_dev(C.byref(N), self._stream()), 'nrst_krige')
                        self._count('launches', 2)
                        continue
                    # estimation variance: A^-1 of the systems, in slices that bound it
                    per = max(1, int((self.aux_limit // 2) // (m * m * 8)))
                    N.ev_out = ev_out.data_ptr()
                    for u0 in range(0, n_grp, per):
                        u1 = min(n_grp, u0 + per)
                        inv = torch.empty((u1 - u0, m, m), dtype=_F64, device=self.device)
                        N.inv, N.u_beg, N.u_end = inv.data_ptr(), u0, u1
                        _lib.check(lib.spx_nrst_solve_dev(C.byref(N), self._stream()), 'nrst_solve')
                        _lib.check(lib.spx_nrst_krige_dev(C.byref(N), self._stream()), 'nrst_krige')
                        self._count('launches', 2)
            self.stats['nrst_systems'] = self.stats.get('nrst_systems', 0) + n_grp

    # ---- IDW ------------------------------------------------------------
    def _idw(self, ctx, out, steps, idw_exp):
        """interp/steps.py:293-313 as  (Z0 . W^T) / (M . W^T)  with W = d**-p."""
        if not steps.size:
            return
        n_stn, n_cells = ctx['n_stn'], ctx['n_cells']
        kpad = _pad_up(n_stn, 8)
        grp_of_step = ctx['grp_of_step']
        # distance scale common to every cell (cancels in the ratio)
        bx0, bx1, by0, by1 = ctx['bbox']
        scale = math.hypot(max(bx1, ctx['stn_xs'].max()) - min(bx0, ctx['stn_xs'].min()),
                           max(by1, ctx['stn_ys'].max()) - min(by0, ctx['stn_ys'].min()))
        scale = scale if scale > 0 else 1.0

        max_grps = max(1, int(self.aux_limit // (n_cells * 8)))
        grps_all = np.unique(grp_of_step[steps])
        for b0 in range(0, grps_all.size, max_grps):
            gb = grps_all[b0:b0 + max_grps]
            slot_of = np.full(ctx['n_grps'], -1, dtype=np.int32)
            slot_of[gb] = np.arange(gb.size, dtype=np.int32)
            st = steps[slot_of[grp_of_step[steps]] >= 0]
            # phase A: sum of weights over the available stations of each group
            n_a = gb.size
            coef_a = torch.zeros(_pad_up(n_a, _lib.SPX_BM) * kpad, dtype=_F64, device=self.device)
            d_mask = self._dev(np.where(ctx['grp_mask'][gb], 1.0, np.nan))
            _lib.check(self.lib.spx_pack_rows_dev(
                self._ptr(d_mask), n_stn, None, n_a, n_stn, kpad, 1, self._ptr(coef_a), 0,
                self._stream()), 'pack_rows')
            aux = torch.empty((n_a, n_cells), dtype=_F64, device=self.device)
            d_slots = torch.arange(n_a, dtype=_I32, device=self.device)
            self._gemm(ctx, coef=coef_a, n_rows=n_a, kpad=kpad, n_border=0, gen=_lib.GEN_IDW,
                       epi=_lib.EPI_AUX, row_dst=d_slots, aux=aux, idw_exp=idw_exp,
                       dist_scale=scale)
            # phase B: data rows, divided by their group's row
            n_b = st.size
            coef_b = torch.zeros(_pad_up(n_b, _lib.SPX_BM) * kpad, dtype=_F64, device=self.device)
            d_st = self._dev(st.astype(np.int32))
            _lib.check(self.lib.spx_pack_rows_dev(
                self._ptr(ctx['d_data0']), n_stn, self._ptr(d_st), n_b, n_stn, kpad, 0,
                self._ptr(coef_b), 0, self._stream()), 'pack_rows')
            d_row_aux = self._dev(slot_of[grp_of_step[st]])
            self._gemm(ctx, coef=coef_b, n_rows=n_b, kpad=kpad, n_border=0, gen=_lib.GEN_IDW,
                       epi=_lib.EPI_FIELD_DIV, row_dst=d_st, row_aux=d_row_aux, out=out, aux=aux,
                       idw_exp=idw_exp, dist_scale=scale)
            self._count('launches', 2)   # the two pack_rows launches

    # ---- kriging --------------------------------------------------------
    def _krige(self, ctx, out, kind_name, steps, step_vg, uniq_vgs, drft_arrs, stns_drft,
               problem_steps, force_direct=False, ev_out=None, no_fast=False):
        """OK / SK / EDK in dual form:
        Z[t, i] = rhs_i . A_g^-1 [z_t; 0]."""
        if not steps.size:
            return
        kind = _lib.KRG_KINDS[kind_name]
        n_stn, n_cells = ctx['n_stn'], ctx['n_cells']
        n_drifts = 0 if kind != 2 else int(stns_drft.shape[1])
        n_border = {0: 1, 1: 0, 2: 1 + n_drifts}[kind]
        kpad = _pad_up(n_stn + n_border, 8)
        grp_of_step, grp_n = ctx['grp_of_step'], ctx['grp_n']

        K = types.SimpleNamespace(kind=kind, n_drifts=n_drifts, n_border=n_border, kpad=kpad,
                                  uniq_vgs=uniq_vgs, keep={})
        K.d_cell_drift = K.d_stn_drift = None
        bad_cells = np.zeros(0, dtype=np.int64)
        drft = None
        if kind == 2:
            drft = np.ascontiguousarray(drft_arrs, dtype=np.float64)
            assert drft.shape == (n_drifts, n_cells)
            K.d_cell_drift = self._dev(drft)
            K.d_stn_drift = self._dev(np.ascontiguousarray(stns_drft, dtype=np.float64))
            ctx['stns_drft_bytes'] = np.ascontiguousarray(stns_drft, dtype=np.float64).tobytes()
            bad_cells = np.where(np.isnan(drft).any(axis=0))[0]

        if (self.native_plan and self.downdate and kind != 1 and not force_direct and
                not no_fast and ev_out is None and not bad_cells.size):
            fn = self._krige_fast(ctx, out, kind_name, K, steps, step_vg, uniq_vgs, drft,
                                  drft_arrs, stns_drft, problem_steps)
            if fn is not NotImplemented:
                return fn

        # station lists per group (ascending station index = reference order); the
        # pseudo-group n_grps holds every station (the "full" system of the downdate)
        n_grps = ctx['n_grps']
        grps_used = np.unique(grp_of_step[steps])
        stn_off = np.zeros(n_grps + 1, dtype=np.int64)
        stn_off[grps_used] = np.concatenate([[0], np.cumsum(grp_n[grps_used])])[:-1]
        stn_off[n_grps] = int(grp_n[grps_used].sum())
        K.stn_off = stn_off
        K.full_grp = n_grps
        rows_sel = np.concatenate([grps_used, [n_grps]]).astype(np.int32)
        K.d_stn_list = self._mask_lists(ctx, rows_sel, stn_off[rows_sel],
                                        int(stn_off[n_grps]) + n_stn, want=1)

        # systems = distinct (group, variogram) pairs among the steps
        pair = grp_of_step[steps].astype(np.int64) * len(uniq_vgs) + step_vg[steps]
        upair, sys_of_row = np.unique(pair, return_inverse=True)
        K.sys_grp = (upair // len(uniq_vgs)).astype(np.int32)
        K.sys_vg = (upair % len(uniq_vgs)).astype(np.int32)
        n_sys = upair.size
        K.sys_n = grp_n[K.sys_grp].astype(np.int32)

        # coefficient rows: steps ordered by variogram, one SPX_BM-aligned segment each
        order = np.argsort(step_vg[steps], kind='stable')
        K.steps_o = steps[order]
        K.sys_o = sys_of_row[order]
        vg_o = step_vg[K.steps_o]
        seg_vgs, seg_first, seg_cnt = np.unique(vg_o, return_index=True, return_counts=True)
        seg_row0 = np.zeros(seg_vgs.size, dtype=np.int64)
        acc = 0
        for k in range(seg_vgs.size):
            seg_row0[k] = acc
            acc += _pad_up(seg_cnt[k], _lib.SPX_BM)
        total_rows = acc
        # Many variograms with few steps each (per-step variogram series): the
        # contraction would regenerate its operand tile per variogram; use the
        # per-row-variogram estimator (row-major coefficients) instead.
        K.local = None
        if self.local_support and ev_out is None and n_drifts <= 4:
            K.local = self._local_plan(ctx, [uniq_vgs[int(v)] for v in seg_vgs])
        mv_smem = ((n_stn + n_border) * 64 + 4 * (n_stn + n_border) + 1024) * 8 + 4096
        K.use_mv = bool(K.local is None and self.multivg and seg_vgs.size >= 8
                        and K.steps_o.size / seg_vgs.size < 32 and mv_smem <= 220 * 1024)
        K.row_major = K.use_mv or (K.local is not None)
        K.row_of = np.empty(K.steps_o.size, dtype=np.int64)
        if K.row_major:
            K.row_of[:] = np.arange(K.steps_o.size)
            total_rows = K.steps_o.size
        else:
            for k in range(seg_vgs.size):
                K.row_of[seg_first[k]:seg_first[k] + seg_cnt[k]] = (
                    seg_row0[k] + np.arange(seg_cnt[k]))
        K.coef = torch.zeros(total_rows * kpad, dtype=_F64, device=self.device)
        row_dst_np = np.full(total_rows, -1, dtype=np.int32)
        row_dst_np[K.row_of] = K.steps_o
        d_row_dst = self._dev(row_dst_np)
        K.rows_by_sys = np.argsort(K.sys_o, kind='stable')
        cnt = np.bincount(K.sys_o, minlength=n_sys)
        K.sys_row_beg = np.concatenate([[0], np.cumsum(cnt)])

        K.d_vgs = self._dev(_lib.vgs_to_numpy(uniq_vgs).view(np.uint8))

        rhs_bound = self._rhs_bound(ctx, K, drft)
        K.rhs_bound = rhs_bound

        # ---- solve: downdated where it pays, direct LU otherwise ----------
        # Everything below is only QUEUED on the stream; residuals / info flags are
        # read back after the contraction has been launched so that the host never
        # idles the GPU (the rare unhealthy cases are then redone).
        resid = np.full(n_sys, np.inf)
        singular = np.zeros(n_sys, dtype=bool)
        vg_cnt = np.bincount(K.sys_vg, minlength=len(uniq_vgs))
        r_sys = n_stn - K.sys_n
        use_dd = np.zeros(n_sys, dtype=bool)
        K.ev_out = ev_out
        if self.downdate and kind != 1 and not force_direct and ev_out is None:
            max_r = self.lib.spx_krige_downdate_max_r()
            use_dd = (vg_cnt[K.sys_vg] >= self.downdate_min_systems) & (r_sys <= max_r)
        pending = []
        if use_dd.any():
            with self._phase('solve_downdate'):
                pending += self._solve_downdate(ctx, K, np.where(use_dd)[0])
        direct = np.where(~use_dd)[0]
        if direct.size:
            with self._phase('solve_direct'):
                pending += self._solve_direct(ctx, K, direct, want_resid=(kind != 1))

        K.flags_event = torch.cuda.Event()
        K.flags_event.record(self._main_stream())

        self._estimate(ctx, K, out, kind, (seg_vgs, seg_first, seg_cnt, seg_row0), d_row_dst,
                       step_vg)

        if ev_out is not None:
            self._est_vars(ctx, K, ev_out)

        # ---- fallbacks that need no host decision (steps.py:418-426) -------
        if bad_cells.size:
            # NaN drift at a cell -> sum(lambda) is NaN -> NNB there, every step
            gl = np.unique(grp_of_step[K.steps_o])
            slot_of = np.full(n_grps, -1, dtype=np.int32)
            slot_of[gl] = np.arange(gl.size, dtype=np.int32)
            nnb = self._nnb_index(ctx, gl, cells=bad_cells)
            pos = ctx['out_pos'][bad_cells] if ctx['out_pos'] is not None else bad_cells
            self._nnb_gather(ctx, out, nnb, K.steps_o, slot_of[grp_of_step[K.steps_o]],
                             n_cells=int(bad_cells.size), d_pos=self._dev(pos.astype(np.int32)))

        self.stats['n_systems'] = self.stats.get('n_systems', 0) + n_sys
        self.stats['n_downdated'] = self.stats.get('n_downdated', 0) + int(use_dd.sum())
        flags_event = K.flags_event

        def deferred():
            """Read the health flags (already in pinned memory), then handle the
            rare unhealthy systems."""
            flags_event.synchronize()
            dd_failed = False
            for fin in pending:
                dd_failed |= fin(resid, singular)
            if dd_failed:
                # an unhealthy full system or r x r block: redo without downdating
                self._count('downdate_redo')
                fn = self._krige(ctx, out, kind_name, steps, step_vg, uniq_vgs, drft_arrs,
                                 stns_drft, problem_steps, force_direct=True, ev_out=ev_out)
                if fn is not None:
                    fn()
                return
            with np.errstate(invalid='ignore'):
                dev = resid * rhs_bound[K.sys_vg]
            flagged = (kind == 1) | singular | ~(dev <= self.lambda_tol)
            for sid in [k_ for k_ in K.keep if not flagged[k_]]:
                del K.keep[sid]
            self.stats['n_flagged'] = self.stats.get('n_flagged', 0) + int(flagged.sum())
            if flagged.any():
                # flagged systems need their factors: (re)do the downdated ones directly
                redo = np.where(flagged & use_dd)[0]
                if redo.size:
                    fins = self._solve_direct(ctx, K, redo, want_resid=False)
                    torch.cuda.current_stream(self.device).synchronize()
                    for fin in fins:
                        fin(resid, singular)
                self._krige_flagged(ctx, out, K, flagged, singular, bad_cells, problem_steps)
            K.keep.clear()

        return deferred

    def _ginv_key(self, ctx, K):
        return (K.kind, K.n_drifts, ctx['min_vg_val'], ctx['stn_xs'].tobytes(),
                ctx['stn_ys'].tobytes(),
                None if K.d_stn_drift is None else ctx['stns_drft_bytes'])

    def _rhs_bound(self, ctx, K, drft):
        """Bound for the sum(lambda) screening: |rhs| <= max(vg bound, 1, |drift|)."""
        bx0, bx1, by0, by1 = ctx['bbox']
        max_dist = math.hypot(max(bx1, ctx['stn_xs'].max()) - min(bx0, ctx['stn_xs'].min()),
                              max(by1, ctx['stn_ys'].max()) - min(by0, ctx['stn_ys'].min()))
        rhs_bound = np.array([max(1.0, vg_abs_bound(v, max_dist)) for v in K.uniq_vgs])
        if K.kind == 2:
            with np.errstate(invalid='ignore'):
                dmax = np.nanmax(np.abs(drft)) if np.isfinite(drft).any() else 1.0
            rhs_bound = np.maximum(rhs_bound, dmax)
        return rhs_bound

    # ---- one native call per chunk (csrc/spx_chunk.cu) ---------------------------
    def _fast_job(self, ctx, vg_s, min_var_thr, n_steps):
        """The native job (ring of pre-allocated slots, solve stream, events) for this
        station set / variogram / grid, created on first use.  None when the chunk
        does not qualify: the full-system inverse is not cached yet (the general path
        builds it), nugget-only variogram, ..."""
        if check_full_nuggetness(vg_s, ctx['min_vg_val']):
            return None
        n_stn, n_cells = ctx['n_stn'], ctx['n_cells']
        n_border = 1
        kpad = _pad_up(n_stn + n_border, 8)
        K = types.SimpleNamespace(kind=0, n_drifts=0, n_border=n_border, kpad=kpad,
                                  uniq_vgs=[vg_s], d_stn_drift=None)
        gkey = (self._ginv_key(ctx, K), vg_s)
        ginv = self._ginv_cache.get(gkey)
        if ginv is None and not (self.sparse_solve and self.local_support):
            return None           # no sparse form either: the general path builds the inverse
        jkey = (gkey, ctx['geom_key'], ctx['out_f64'], ctx['has_lo'], ctx['has_hi'], ctx['lo'],
                ctx['hi'], min_var_thr, self.local_support, self.local_max_near,
                self.local_tiles, self.lambda_tol, self.downdate_min_systems, self.solve_stream,
                self.sparse_solve)
        job = self._fast_jobs.get(jkey)
        if job is not None and job['max_steps'] >= n_steps and (
                job['ginv'] is ginv or job['cfg'].sparse.n_comp > 0):
            return job
        if ginv is None and jkey in self._fast_jobs_none:
            return None           # the sparse form does not apply here (checked before)
        if job is not None:
            self._fast_job_close(jkey)
        lib = self.lib
        cfg = _lib.spx_fast_cfg()
        cfg.n_stn, cfg.n_border, cfg.kpad = n_stn, n_border, kpad
        cfg.max_steps = int(n_steps)
        cfg.n_slots = int(self.fast_slots)
        cfg.min_systems = int(self.downdate_min_systems)
        cfg.min_var_thr = float(min_var_thr)
        cfg.ginv = ginv.data_ptr() if ginv is not None else None
        cfg.lambda_bound = float(self._rhs_bound(ctx, K, None)[0])
        cfg.lambda_tol = float(self.lambda_tol)
        cfg.profile = 1
        cfg.solve_stream = int(bool(self.solve_stream))
        keep = [ginv, ctx['d_stn_x'], ctx['d_stn_y'], ctx['d_cell_x'], ctx['d_cell_y'],
                ctx['d_pos']]
        local = self._local_plan(ctx, [vg_s]) if self.local_support else None
        if local is not None:
            nbr = self._local_neighbours(ctx, K, vg_s, local[0])
            keep.append(nbr)
            L = nbr['struct']
            L.kpad, L.n_stn, L.n_drifts = kpad, n_stn, 0
            L.cell_drift = None
            L.out_ld = ctx['fld_size']
            L.out_f64 = ctx['out_f64']
            L.cell_pos = ctx['d_pos'].data_ptr() if ctx['d_pos'] is not None else None
            L.has_lo, L.has_hi, L.lo, L.hi = ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi']
            L.rows_all_valid = 1
            cfg.estimator = 0
            cfg.want_coef_t = int(self._local_wants_coef_t(ctx, 0))
            tot = local[0][1]
            cfg.base_f = float(tot if tot > ctx['min_vg_val'] else 0.0)
            cfg.local = L
            sp = self._sparse_cov(ctx, vg_s, local[0][0], cfg.base_f) if self.sparse_solve else None
            if sp is not None:
                keep.append(sp)
                cfg.sparse = sp['struct']
        else:
            g = _lib.spx_gemm()
            g.kpad, g.n_stn, g.n_border = kpad, n_stn, n_border
            g.stn_x, g.stn_y = ctx['d_stn_x'].data_ptr(), ctx['d_stn_y'].data_ptr()
            g.cell_x, g.cell_y = ctx['d_cell_x'].data_ptr(), ctx['d_cell_y'].data_ptr()
            g.n_cells = n_cells
            g.cell_drift = None
            g.gen, g.covar_flag = _lib.GEN_VG, 0
            g.vg = _lib.make_vg(vg_s)
            g.min_vg_val = ctx['min_vg_val']
            g.idw_exp, g.dist_scale = 0.0, 1.0
            g.epi = _lib.EPI_FIELD
            g.row_aux = None
            g.out_ld = ctx['fld_size']
            g.out_f64 = ctx['out_f64']
            g.cell_pos = ctx['d_pos'].data_ptr() if ctx['d_pos'] is not None else None
            g.aux = None
            g.has_lo, g.has_hi, g.lo, g.hi = ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi']
            g.quad_slot = 0
            cfg.estimator = 1
            cfg.want_coef_t = 0
            cfg.gemm = g
        if ginv is None and cfg.sparse.n_comp == 0:
            # neither a cached inverse of the full station system nor the sparse form
            self._fast_jobs_none.add(jkey)
            return None
        d_bytes = int(lib.spx_fast_slot_bytes(C.byref(cfg), 0))
        h_bytes = int(lib.spx_fast_slot_bytes(C.byref(cfg), 1))
        if d_bytes <= 0 or h_bytes <= 0:
            return None
        self._sync_uploads()
        # tables built just now (first chunk of a job) are read by the solve stream
        torch.cuda.current_stream(self.device).synchronize()
        d_arena = torch.empty(cfg.n_slots * d_bytes, dtype=torch.uint8, device=self.device)
        h_arena = torch.empty(cfg.n_slots * h_bytes, dtype=torch.uint8, pin_memory=True)
        handle = C.c_void_p()
        _lib.check(lib.spx_fast_create(C.byref(cfg), C.c_void_p(d_arena.data_ptr()),
                                       C.c_void_p(h_arena.data_ptr()), C.byref(handle)),
                   'fast_create')
        W = (n_stn + 63) // 64
        job = dict(handle=handle, cfg=cfg, keep=keep, d_arena=d_arena, h_arena=h_arena,
                   max_steps=int(n_steps), ginv=ginv, estimator=int(cfg.estimator), W=W,
                   next_slot=0, id=self._next_job_id(), refs=0, retired=False, unchecked={},
                   res=_lib.spx_fast_result())
        while len(self._fast_jobs) >= 2:
            self._fast_job_close(next(iter(self._fast_jobs)))
        self._fast_jobs[jkey] = job
        return job

    def _sparse_cov(self, ctx, vg_s, R, base_f):
        """Connected components of the graph "stations closer than the variogram range" and
        their covariance blocks F - vg(d) on the device (spx_sparse_cov), or None when a
        component has more than SPX_SPARSE_MAX_COMP stations: beyond the range the variogram
        is the constant F, so the ordinary-kriging matrix is F 11' - C with C block diagonal
        over these components and a time step is solved in O(n_stn)
        (spx_krige_sparse_ok_dev) instead of a dense factorisation -- same system, same
        solution to rounding."""
        n = ctx['n_stn']
        lab = station_clusters(ctx['stn_xs'], ctx['stn_ys'], R, _lib.SPX_SPARSE_MAX_COMP)
        if lab is None:
            return None
        n_comp = int(lab.max()) + 1
        sizes = np.bincount(lab, minlength=n_comp)
        # components by size (singles first: the kernel's lanes then share their code path),
        # members of a component by station index
        rank = np.empty(n_comp, dtype=np.int64)
        rank[np.argsort(sizes, kind='stable')] = np.arange(n_comp)
        lab = rank[lab]
        sizes = np.bincount(lab, minlength=n_comp)
        order = np.argsort(lab, kind='stable').astype(np.int32)       # members by component
        comp_off = np.zeros(n_comp + 1, dtype=np.int32)
        np.cumsum(sizes, out=comp_off[1:])
        blk_off = np.zeros(n_comp, dtype=np.int64)
        np.cumsum((sizes.astype(np.int64) ** 2)[:-1], out=blk_off[1:])
        n_blk = int((sizes.astype(np.int64) ** 2).sum())
        d_off = torch.from_numpy(comp_off).to(self.device)
        d_stn = torch.from_numpy(order).to(self.device)
        d_boff = torch.from_numpy(blk_off).to(self.device)
        d_blk = torch.empty(n_blk, dtype=_F64, device=self.device)
        sp = _lib.spx_sparse_cov()
        sp.n_comp, sp.max_size = int(n_comp), int(sizes.max())
        sp.n_single = int((sizes == 1).sum())
        sp.comp_off, sp.comp_stn = d_off.data_ptr(), d_stn.data_ptr()
        sp.blk_off, sp.blk = d_boff.data_ptr(), d_blk.data_ptr()
        vg = _lib.make_vg(vg_s)
        _lib.check(self.lib.spx_sparse_cov_blocks_dev(
            C.c_void_p(ctx['d_stn_x'].data_ptr()), C.c_void_p(ctx['d_stn_y'].data_ptr()),
            C.byref(sp), C.byref(vg), float(ctx['min_vg_val']), float(base_f), self._stream()),
            'sparse_cov_blocks')
        self._count('launches')
        self._count('sparse_cov_jobs')
        return dict(struct=sp, tensors=(d_off, d_stn, d_boff, d_blk), n_comp=int(n_comp),
                    max_size=int(sizes.max()))

    def _next_job_id(self):
        self._fast_job_seq = getattr(self, '_fast_job_seq', 0) + 1
        return self._fast_job_seq

    def _fast_job_close(self, jkey):
        """Drop a job from the cache; it is destroyed once no pending chunk refers to it
        (a chunk's health check reads the job's slot)."""
        job = self._fast_jobs.pop(jkey, None)
        if job is not None:
            job['retired'] = True
            self._fast_job_release(job, 0)

    def _fast_job_release(self, job, n=1):
        job['refs'] -= n
        if job.get('retired') and job['refs'] <= 0 and job['handle'] is not None:
            self._fast_collect(job, all_slots=True)
            _lib.check(self.lib.spx_fast_destroy(job['handle']), 'fast_destroy')
            job['handle'] = None

    def close(self):
        """Release the native jobs (their streams and events)."""
        for jkey in list(self._fast_jobs):
            self._fast_job_close(jkey)

    def _fast_collect(self, job, slot=None, all_slots=False):
        """Move the measured estimate times of finished slots into kernel_events
        (profile_gemm): a slot's events are re-recorded when the ring comes round."""
        if job['handle'] is None:
            return
        slots = range(job['cfg'].n_slots) if all_slots else [slot]
        for k in slots:
            rec = self._fast_prof.pop((job['id'], k), None)
            if rec is None:
                continue
            ms, ms_solve = C.c_float(0.0), C.c_float(0.0)
            _lib.check(self.lib.spx_fast_times(job['handle'], k, C.byref(ms), C.byref(ms_solve)),
                       'fast_times')
            self.kernel_events.append(rec + (float(ms.value), None))
            self.solve_ms.append(float(ms_solve.value))

    def collect_profile(self):
        """Finish kernel_events (synchronises the estimate launches still in flight)."""
        for job in self._fast_jobs.values():
            self._fast_collect(job, all_slots=True)

    def _fast_submit(self, ctx, data, vg_s, min_var_thr, tdtype, label):
        """Queue the whole kriging label of the chunk with one native call; None if
        the chunk does not qualify (nothing queued then)."""
        n_steps, n_stn = ctx['n_steps'], ctx['n_stn']
        job = self._fast_job(ctx, vg_s, min_var_thr, n_steps)
        if job is None:
            return None
        if job['estimator'] == 0 and not self.local_support:
            return None
        early = job['unchecked'].get(job['next_slot'])
        if early is not None and early['run'] is not None:
            early['run']()   # the ring came round: read that chunk's health flags first
        if self.profile_gemm or self._fast_prof:
            self._fast_collect(job, slot=job['next_slot'])
        if ctx['out_pos'] is None:
            out = torch.empty((n_steps, ctx['fld_size']), dtype=tdtype, device=self.device)
        else:
            out = torch.full((n_steps, ctx['fld_size']), float('nan'), dtype=tdtype,
                             device=self.device)
        ints = np.empty((4, n_steps), dtype=np.int32)    # grp_of_step, grp_first, grp_n, n_avail
        step_flag = np.empty(n_steps, dtype=np.uint8)
        grp_bits = np.empty((n_steps, job['W']), dtype=np.uint64)
        res = _lib.spx_fast_result()
        self._sync_uploads()
        self._main_stream()
        _lib.check(self.lib.spx_fast_submit(
            job['handle'], data.ctypes.data, n_steps, data.strides[0] // 8,
            C.c_void_p(out.data_ptr()), self._main_handle, ints[0].ctypes.data,
            ints[1].ctypes.data, ints[2].ctypes.data, ints[3].ctypes.data, step_flag.ctypes.data,
            grp_bits.ctypes.data, C.byref(res)), 'fast_submit')
        g = int(res.n_grps)
        groups = (ints[0], grp_bits[:g], ints[2, :g].astype(np.int64), ints[1, :g],
                  ints[3].astype(np.int64), step_flag.view(np.bool_))
        if res.status != 0:
            # not eligible (too few systems, too many missing stations): the grouping is
            # still good, the general path takes the label
            self._count('fast_not_eligible')
            return dict(label=None, groups=groups, res=res, out=None, job=job)
        job['next_slot'] = (int(res.slot) + 1) % job['cfg'].n_slots
        job['refs'] += 1                 # released by the chunk's deferred health check
        hm = self.stats.setdefault('fast_host_ms', [0.0] * 6)
        for i in range(6):
            hm[i] += res.host_ms[i]
        self._count('launches', int(res.launches))
        self._count('native_submits')
        self._count('n_systems', int(res.n_sys))
        self._count('n_downdated', int(res.n_sys))
        self._count('ginv_cache_hits')
        self.h2d_bytes += int(res.h2d_bytes)
        n_cells = ctx['n_cells']
        esz = 8 if ctx['out_f64'] else 4
        if job['estimator'] == 0:
            self._count('local_rows', int(res.n_krige))
            self._count('local_cache_hits')
            rec = ('k_estimate_local', 'hbm', float(int(res.n_krige) * n_cells * esz))
        else:
            flop = 2.0 * int(res.n_krige) * job['cfg'].kpad * n_cells
            self._count('gemm_launches')
            self._count('gemm_flop', int(flop))
            rec = ('k_estimate_gemm', 'tensor', flop)
        if self.profile_gemm:
            self._fast_prof[(job['id'], int(res.slot))] = rec
        return dict(label=label, groups=groups, res=res, out=out, job=job, vg=vg_s)

    def _fast_deferred(self, ctx, fast, out, krige_mask, problem_steps):
        job, slot = fast['job'], int(fast['res'].slot)
        # state <-> deferred is a reference cycle only until the check has run (it is cut
        # there): the closure keeps the chunk's 5 GB field alive, and a cycle would leave its
        # release to the garbage collector
        state = {'done': False, 'run': None}

        def deferred():
            """Health flags of the slot (mapped host memory): an unhealthy elimination or
            a system whose weights may not sum to one sends the label through the general
            path, which knows every fallback of the reference.  Runs once: from
            PendingChunk.result(), or earlier if the ring comes round to the slot."""
            if state['done']:
                return
            state['done'] = True
            state['run'] = None
            if job['unchecked'].get(slot) is state:
                del job['unchecked'][slot]
            verdict = C.c_int32(0)
            try:
                _lib.check(self.lib.spx_fast_check(job['handle'], slot, C.byref(verdict)),
                           'fast_check')
            finally:
                self._fast_job_release(job)
            if verdict.value == 0:
                self.stats['n_flagged'] = self.stats.get('n_flagged', 0)
                return
            unhealthy = verdict.value == 1
            self._count('downdate_redo' if unhealthy else 'fast_path_redo')
            steps = np.where(krige_mask)[0]
            fn = self._krige(ctx, out, 'OK', steps, np.zeros(ctx['n_steps'], dtype=np.int32),
                             [fast['vg']], None, None, problem_steps, force_direct=unhealthy,
                             no_fast=True)
            if fn is not None:
                fn()

        state['run'] = deferred
        job['unchecked'][slot] = state
        return deferred

    def _krige_fast(self, ctx, out, kind_name, K, steps, step_vg, uniq_vgs, drft, drft_arrs,
                    stns_drft, problem_steps):
        """The common case of _krige -- every system downdated from a cached full-system
        inverse -- with the index work done by the native planner: per variogram ONE
        planner call, one small upload, then station lists / Bt on the device, Ut = Bt G,
        the downdate kernel (which also emits the transposed coefficients and base values
        of the local estimator) and the estimate stage.  Returns NotImplemented when a
        precondition fails (the general path then runs); anything unhealthy found later is
        redone by the general path as well."""
        lib = self.lib
        n_stn, n_cells = ctx['n_stn'], ctx['n_cells']
        kpad, n_border, n_drifts, kind = K.kpad, K.n_border, K.n_drifts, K.kind
        M = n_stn + n_border
        if not self.ginv_cache:
            return NotImplemented
        steps = np.ascontiguousarray(steps, dtype=np.int32)
        # steps ordered by variogram (stable), one segment per variogram
        if len(uniq_vgs) == 1:
            steps_o = steps
            seg_vgs = np.zeros(1, dtype=np.int64)
            seg_first = np.zeros(1, dtype=np.int64)
            seg_cnt = np.array([steps.size], dtype=np.int64)
        else:
            vg_of = step_vg[steps]
            order = np.argsort(vg_of, kind='stable')
            steps_o = np.ascontiguousarray(steps[order])
            seg_vgs, seg_first, seg_cnt = np.unique(vg_of[order], return_index=True,
                                                    return_counts=True)
        base_key = self._ginv_key(ctx, K)
        ginvs = []
        for v in seg_vgs:
            hit = self._ginv_cache.get((base_key, uniq_vgs[int(v)]))
            if hit is None:
                return NotImplemented          # the general path builds and caches it
            ginvs.append(hit)
        # estimator choice exactly as in the general path
        K.local = None
        if self.local_support and n_drifts <= 4:
            K.local = self._local_plan(ctx, [uniq_vgs[int(v)] for v in seg_vgs])
        mv_smem = ((n_stn + n_border) * 64 + 4 * (n_stn + n_border) + 1024) * 8 + 4096
        K.use_mv = bool(K.local is None and self.multivg and seg_vgs.size >= 8
                        and steps_o.size / seg_vgs.size < 32 and mv_smem <= 220 * 1024)
        K.row_major = K.use_mv or (K.local is not None)
        seg_row0 = np.zeros(seg_vgs.size, dtype=np.int64)
        if K.row_major:
            row_of = np.arange(steps_o.size, dtype=np.int64)
            total_rows = int(steps_o.size)
            row_dst_np = steps_o
        else:
            row_of = np.empty(steps_o.size, dtype=np.int64)
            acc = 0
            for k in range(seg_vgs.size):
                seg_row0[k] = acc
                row_of[seg_first[k]:seg_first[k] + seg_cnt[k]] = acc + np.arange(seg_cnt[k])
                acc += _pad_up(seg_cnt[k], _lib.SPX_BM)
            total_rows = acc
            row_dst_np = np.full(total_rows, -1, dtype=np.int32)
            row_dst_np[row_of] = steps_o
        # plans first: eligibility (every system downdated by the register kernel) is
        # known before anything is queued
        grp_of_step = ctx['grp_of_step']
        grp_n32 = ctx.get('grp_n32')
        if grp_n32 is None:
            grp_n32 = ctx['grp_n32'] = np.ascontiguousarray(ctx['grp_n'], dtype=np.int32)
        reg_max = lib.spx_krige_downdate_reg_max_r()
        plans = []
        for k in range(seg_vgs.size):
            r0, r1 = int(seg_first[k]), int(seg_first[k] + seg_cnt[k])
            st_k = steps_o[r0:r1]
            rows_k = row_of[r0:r1]
            hb = np.empty(lib.spx_downdate_plan_host_bytes(r1 - r0), dtype=np.uint8)
            plan = _lib.spx_dd_plan()
            _lib.check(lib.spx_downdate_plan_host(
                grp_of_step.ctypes.data, grp_n32.ctypes.data, ctx['n_grps'], n_stn,
                st_k.ctypes.data, rows_k.ctypes.data, r1 - r0, hb.ctypes.data, hb.size,
                C.byref(plan)), 'downdate_plan_host')
            if plan.n_sys < self.downdate_min_systems or plan.max_r > reg_max:
                return NotImplemented
            plans.append((plan, hb))

        K.steps_o = steps_o
        K.coef = torch.zeros(total_rows * kpad, dtype=_F64, device=self.device)
        d_row_dst = self._dev(np.ascontiguousarray(row_dst_np, dtype=np.int32))
        K.d_vgs = self._dev(_lib.vgs_to_numpy(uniq_vgs).view(np.uint8)) if K.use_mv else None
        rhs_bound = self._rhs_bound(ctx, K, drft)
        want_t = K.local is not None and self._local_wants_coef_t(ctx, n_drifts)
        d_data = ctx['d_data']
        fused = []
        checks = []
        n_sys_total = 0
        with self._phase('solve_downdate'):
            for k, (plan, hb) in enumerate(plans):
                r0, r1 = int(seg_first[k]), int(seg_first[k] + seg_cnt[k])
                n_sys, n_data, n_rhs = plan.n_sys, plan.n_data, plan.n_rhs
                n_sys_total += n_sys
                d_plan = self._arena_take(int(plan.n_bytes))
                _lib.check(lib.spx_upload_dev(C.c_void_p(d_plan.data_ptr()),
                                              C.c_void_p(hb.ctypes.data),
                                              int(plan.n_upload_bytes), self._h2d_handle), 'upload')
                self._h2d_dirty = True
                self.h2d_bytes += int(plan.n_upload_bytes)
                pb = d_plan.data_ptr()
                _lib.check(lib.spx_avail_lists_dev(
                    self._ptr(d_data), n_stn, n_stn, C.c_void_p(pb + plan.off_bt_step + 4 * n_data),
                    n_sys, C.c_void_p(pb + plan.off_sys_stn_off), C.c_void_p(pb + plan.off_stn_list),
                    C.c_void_p(pb + plan.off_sys_miss_off), C.c_void_p(pb + plan.off_miss_list),
                    self._stream()), 'avail_lists')
                Bt = torch.empty((n_rhs, M), dtype=_F64, device=self.device)
                _lib.check(lib.spx_build_bt_dev(
                    self._ptr(d_data), n_stn, n_stn, C.c_void_p(pb + plan.off_bt_step), n_rhs,
                    n_data, n_border, self._ptr(Bt), self._stream()), 'build_bt')
                Ut = torch.matmul(Bt, ginvs[k])
                self._count('launches', 3)
                # resid [n_rhs] f64 (zero-initialised) followed by info [n_sys] i32
                flags = torch.zeros(n_rhs + (n_sys + 1) // 2, dtype=_F64, device=self.device)
                D = _lib.spx_downdate()
                D.n_sys, D.n_stn, D.n_border, D.max_r = n_sys, n_stn, n_border, int(plan.max_r)
                D.ginv = ginvs[k].data_ptr()
                D.sys_r = pb + plan.off_sys_r
                D.sys_miss_off = pb + plan.off_sys_miss_off
                D.miss_list = pb + plan.off_miss_list
                D.sys_n = pb + plan.off_sys_n
                D.sys_stn_off = pb + plan.off_sys_stn_off
                D.stn_list = pb + plan.off_stn_list
                D.sys_rhs_off = pb + plan.off_sys_rhs_off
                D.sys_rhs_cnt = pb + plan.off_sys_rhs_cnt
                D.rhs_urow = pb + plan.off_rhs_urow
                D.rhs_row = pb + plan.off_rhs_row
                D.rhs_kind = pb + plan.off_rhs_kind
                D.sys_order = pb + plan.off_sys_order
                D.ut = Ut.data_ptr()
                D.kpad = kpad
                D.coef = K.coef.data_ptr()
                D.coef_row_major = int(K.row_major)
                D.resid = flags.data_ptr()
                D.info = flags.data_ptr() + 8 * n_rhs
                pre = None
                if K.local is not None:
                    # base / coef_t index rows relative to the segment start
                    pre = dict(base=torch.empty(r1 - r0, dtype=_F64, device=self.device))
                    tot = K.local[k][1]            # F of _local_neighbours (variogram form)
                    F = tot if tot > ctx['min_vg_val'] else 0.0
                    D.base = pre['base'].data_ptr() - 8 * r0
                    D.base_f = float(F)
                    if want_t:
                        ld_t = _pad_up(r1 - r0, 4)
                        pre['coef_t'] = torch.empty((kpad, ld_t), dtype=_F64, device=self.device)
                        pre['ld_t'] = ld_t
                        D.coef_t = pre['coef_t'].data_ptr() - 8 * r0
                        D.coef_t_ld = ld_t
                fused.append(pre)
                _lib.check(lib.spx_krige_downdate_dev(C.byref(D), self._stream()), 'downdate')
                self._count('launches')
                self._count('ginv_cache_hits')
                h_flags = self._fetch_async(flags)
                pos_ones = hb[plan.off_pos_ones:plan.off_pos_ones + 8 * n_sys].view(np.int64)
                checks.append((h_flags, pos_ones, n_rhs, n_sys, float(rhs_bound[int(seg_vgs[k])])))

        flags_event = torch.cuda.Event()
        flags_event.record(self._main_stream())
        self._estimate(ctx, K, out, kind, (seg_vgs, seg_first, seg_cnt, seg_row0), d_row_dst,
                       step_vg, fused=fused)
        self.stats['n_systems'] = self.stats.get('n_systems', 0) + n_sys_total
        self.stats['n_downdated'] = self.stats.get('n_downdated', 0) + n_sys_total
        self._count('native_plans', len(plans))

        def deferred():
            """Health flags (already in pinned memory): an unhealthy elimination or a
            system whose weights may not sum to one sends the whole call through the
            general path, which knows every fallback of the reference."""
            flags_event.synchronize()
            unhealthy = False
            flagged = 0
            for h_flags, pos_ones, n_rhs, n_sys, bound in checks:
                hf = h_flags.numpy()
                info = hf[n_rhs:].view(np.int32)[:n_sys]
                unhealthy |= bool((info != 0).any())
                with np.errstate(invalid='ignore'):
                    dev = hf[:n_rhs][pos_ones] * bound
                flagged += int((~(dev <= self.lambda_tol)).sum())
            if unhealthy or flagged:
                self._count('downdate_redo' if unhealthy else 'fast_path_redo')
                fn = self._krige(ctx, out, kind_name, steps, step_vg, uniq_vgs, drft_arrs,
                                 stns_drft, problem_steps, force_direct=unhealthy, no_fast=True)
                if fn is not None:
                    fn()
            else:
                self.stats['n_flagged'] = self.stats.get('n_flagged', 0)

        return deferred

    @staticmethod
    def _local_wants_coef_t(ctx, n_drifts):
        """The streamlined local kernel (transposed coefficients) applies."""
        return (not ctx['out_f64']) and n_drifts == 0 and ctx['d_pos'] is None

    def _estimate(self, ctx, K, out, kind, segs, d_row_dst, step_vg, fused=None):
        """Estimate stage of _krige: coefficient rows -> field rows, one variogram segment
        at a time (local estimator / per-row-variogram estimator / DMMA contraction).
        fused[k]: per-segment dict with 'base' / 'coef_t' / 'ld_t' already written by the
        downdate kernel."""
        seg_vgs, seg_first, seg_cnt, seg_row0 = segs
        kpad, n_border, n_drifts = K.kpad, K.n_border, K.n_drifts
        n_stn, n_cells = ctx['n_stn'], ctx['n_cells']
        if K.local is not None:
            coef2d = K.coef.view(-1, kpad)
            # more variograms than the table cache holds (per-step variogram series): the
            # tables are built per variogram into shared buffers sized by the largest range
            transient = None
            if seg_vgs.size > 4:
                transient = {}
                k_max = int(np.argmax([pl[0] for pl in K.local]))
                self._local_neighbours(ctx, K, K.uniq_vgs[int(seg_vgs[k_max])], K.local[k_max],
                                       transient=transient)
            for k in range(seg_vgs.size):
                r0, r1 = int(seg_first[k]), int(seg_first[k] + seg_cnt[k])
                nbr = self._local_neighbours(ctx, K, K.uniq_vgs[int(seg_vgs[k])], K.local[k],
                                             transient=transient)
                pre = fused[k] if fused is not None else None
                if pre is not None and pre.get('base') is not None:
                    base = pre['base']             # written by the downdate kernel
                else:
                    base = nbr['F'] * coef2d[r0:r1, :n_stn].sum(dim=1)
                    if n_border >= 1:
                        base = base + coef2d[r0:r1, n_stn]
                    base = base.contiguous()
                L = nbr['struct']
                L.coef = coef2d[r0:r1].data_ptr()
                L.base = base.data_ptr()
                L.n_rows = r1 - r0
                L.kpad, L.n_stn, L.n_drifts = kpad, n_stn, n_drifts
                L.cell_drift = K.d_cell_drift.data_ptr() if K.d_cell_drift is not None else None
                L.row_dst = d_row_dst[r0:].data_ptr()
                L.out = out.data_ptr()
                L.out_ld = ctx['fld_size']
                L.out_f64 = ctx['out_f64']
                L.cell_pos = ctx['d_pos'].data_ptr() if ctx['d_pos'] is not None else None
                L.has_lo, L.has_hi, L.lo, L.hi = ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi']
                L.rows_all_valid = 1          # row-major rows are exactly the kriged steps
                L.coef_t, L.coef_t_ld = None, 0
                if pre is not None and pre.get('coef_t') is not None:
                    L.coef_t, L.coef_t_ld = pre['coef_t'].data_ptr(), pre['ld_t']
                elif self._local_wants_coef_t(ctx, n_drifts):
                    # transposed copy for the streamlined kernel (5 MB at 1250 x 504)
                    ld_t = _pad_up(r1 - r0, 4)
                    coef_t = torch.zeros((kpad, ld_t), dtype=_F64, device=self.device)
                    coef_t[:, :r1 - r0] = coef2d[r0:r1].t()
                    L.coef_t, L.coef_t_ld = coef_t.data_ptr(), ld_t
                stream = self._stream()
                ev = self._prof_begin()
                _lib.check(self.lib.spx_estimate_local_dev(C.byref(L), stream), 'estimate_local')
                self._prof_end(ev, 'k_estimate_local', 'hbm',
                               (r1 - r0) * n_cells * (8 if ctx['out_f64'] else 4))
                self._count('launches')
                self._count('local_rows', r1 - r0)

        if K.use_mv:
            d_row_vg = self._dev(step_vg[K.steps_o].astype(np.int32))
            g = _lib.spx_multivg()
            g.coef = K.coef.data_ptr()
            g.n_rows = int(K.steps_o.size)
            g.kpad, g.n_stn, g.n_border = kpad, n_stn, n_border
            g.stn_x, g.stn_y = ctx['d_stn_x'].data_ptr(), ctx['d_stn_y'].data_ptr()
            g.cell_x, g.cell_y = ctx['d_cell_x'].data_ptr(), ctx['d_cell_y'].data_ptr()
            g.n_cells = n_cells
            g.cell_drift = K.d_cell_drift.data_ptr() if K.d_cell_drift is not None else None
            g.vgs = K.d_vgs.data_ptr()
            g.row_vg = d_row_vg.data_ptr()
            g.covar_flag = int(kind == 1)
            g.min_vg_val = ctx['min_vg_val']
            g.row_dst = d_row_dst.data_ptr()
            g.out = out.data_ptr()
            g.out_ld = ctx['fld_size']
            g.out_f64 = ctx['out_f64']
            g.cell_pos = ctx['d_pos'].data_ptr() if ctx['d_pos'] is not None else None
            g.has_lo, g.has_hi, g.lo, g.hi = ctx['has_lo'], ctx['has_hi'], ctx['lo'], ctx['hi']
            def _fast(vg_s):       # <= 2 Sph/Lin and <= 2 Exp/Gau terms plus nuggets
                n_poly = n_exp = 0
                for (t, _, _) in _lib.parse_vg_str(vg_s):
                    if t in (2, 4):
                        n_poly += 1
                    elif t in (3, 5):
                        n_exp += 1
                    elif t != 1:
                        return False
                return n_poly <= 2 and n_exp <= 2
            g.all_fast = int(all(_fast(vg_s) for vg_s in K.uniq_vgs))
            with self._phase('multivg'):
                _lib.check(self.lib.spx_estimate_multivg_dev(C.byref(g), self._stream()),
                           'estimate_multivg')
            self._count('launches')
            self._count('multivg_evals', int(K.steps_o.size) * n_stn * n_cells)

        # ---- main contraction: one launch per variogram segment ----------
        for k in range(seg_vgs.size if not K.row_major else 0):
            seg_coef = K.coef[seg_row0[k] * kpad:]
            with self._phase('gemm'):
                self._gemm(ctx, coef=seg_coef, n_rows=int(seg_cnt[k]), kpad=kpad,
                           n_border=n_border, gen=_lib.GEN_VG, epi=_lib.EPI_FIELD,
                           row_dst=d_row_dst[seg_row0[k]:], out=out,
                           vg=_lib.make_vg(K.uniq_vgs[int(seg_vgs[k])]),
                           covar_flag=int(kind == 1), cell_drift=K.d_cell_drift)


    # ---- pseudo-inverse path for numerically singular systems ---------------
    def _pinv_systems(self, ctx, out, K, sids, sv, v, coef_a, row_major=False):
        """np.linalg.pinv semantics (rcond = 1e-15) through a symmetric
        eigendecomposition on the device.  Rewrites the estimates of the steps of
        the listed systems and puts pinv(A) [1; 0] into the rows of ``coef_a`` that
        the caller contracts for the sum(lambda) test.  Rare path (cuSOLVER syevd
        via torch.linalg.eigh)."""
        n_stn, kpad = ctx['n_stn'], K.kpad
        rows_all, cols_all, step_all = [], [], []
        blocks = []
        for sid in sids:
            sid = int(sid)
            n = int(K.sys_n[sid])
            m = n + K.n_border
            T = self._systems_struct(ctx, K, np.array([sid]), K.sys_grp[[sid]], K.sys_vg[[sid]])
            _lib.check(self.lib.spx_krige_assemble_dev(
                C.byref(T.S), self._ptr(K.d_vgs), len(K.uniq_vgs), ctx['min_vg_val'],
                self._stream()), 'assemble')
            self._count('launches')
            A = T.work.view(m, m)
            A = 0.5 * (A + A.T)
            w, V = torch.linalg.eigh(A)
            big = w.abs() > 1e-15 * w.abs().max()
            winv = torch.where(big, 1.0 / torch.where(big, w, torch.ones_like(w)),
                               torch.zeros_like(w))
            pinv = (V * winv[None, :]) @ V.T
            stn = np.where(ctx['grp_mask'][int(K.sys_grp[sid])])[0]
            cols = self._dev(np.concatenate([stn, n_stn + np.arange(K.n_border)]).astype(np.int64))
            ridx = K.rows_by_sys[K.sys_row_beg[sid]:K.sys_row_beg[sid + 1]]
            st = K.steps_o[ridx]
            d_st = self._dev(st.astype(np.int64))
            d_stn = self._dev(stn.astype(np.int64))
            self._sync_uploads()
            B = torch.zeros((st.size + 1, m), dtype=_F64, device=self.device)
            B[:st.size, :n] = ctx['d_data'].index_select(0, d_st).index_select(1, d_stn)
            B[st.size, :n] = 1.0
            Csol = B @ pinv                                   # pinv is symmetric
            full = torch.zeros((st.size + 1, kpad), dtype=_F64, device=self.device)
            full[:, cols] = Csol
            blocks.append((full, st, int(np.searchsorted(sv, sid))))
            self._count('pinv_systems')
        # data rows of all listed systems: one packed segment, one contraction
        n_rows = sum(b[1].size for b in blocks)
        if n_rows:
            dense = torch.cat([b[0][:-1] for b in blocks], dim=0).contiguous()
            coef_p = torch.zeros(_pad_up(n_rows, _lib.SPX_BM) * kpad, dtype=_F64,
                                 device=self.device)
            _lib.check(self.lib.spx_pack_rows_dev(
                self._ptr(dense), kpad, None, n_rows, kpad, kpad, 0, self._ptr(coef_p), 0,
                self._stream()), 'pack_rows')
            d_dst = self._dev(np.concatenate([b[1] for b in blocks]).astype(np.int32))
            self._gemm(ctx, coef=coef_p, n_rows=n_rows, kpad=kpad, n_border=K.n_border,
                       gen=_lib.GEN_VG, epi=_lib.EPI_FIELD, row_dst=d_dst, out=out,
                       vg=_lib.make_vg(K.uniq_vgs[v]), covar_flag=int(K.kind == 1),
                       cell_drift=K.d_cell_drift)
        # ones-vector rows into the caller's segment
        for full, st, row in blocks:
            if row_major:
                coef_a.view(-1, kpad)[row] = full[-1]
                continue
            _lib.check(self.lib.spx_pack_rows_dev(
                self._ptr(full[-1:].contiguous()), kpad, None, 1, kpad, kpad, 0,
                self._ptr(coef_a), row, self._stream()), 'pack_rows')
        self._count('launches', 1 + len(blocks))

    # ---- OK estimation variance (weights form) -----------------------------
    def _est_vars(self, ctx, K, ev_out):
        """est_var = rhs' A^-1 rhs + lambda[n] per (system, cell)
        (interp/steps.py:431-434; the Lagrange term counted twice is the
        reference's, quirk Q8), spread to the steps of the system.  The inverse's
        rows are the coefficient rows of one more contraction whose epilogue
        contracts with the resident right-hand-side tile (SPX_EPI_QUADFORM)."""
        n_cells, n_stn = ctx['n_cells'], ctx['n_stn']
        n_sys = K.sys_n.size
        max_slots = max(1, int(self.aux_limit // (n_cells * 8)))
        for c0 in range(0, n_sys, max_slots):
            ids = np.arange(c0, min(n_sys, c0 + max_slots))
            aux = torch.empty((ids.size, n_cells), dtype=_F64, device=self.device)
            for slot, sid in enumerate(ids):
                T, k = K.keep[int(sid)]
                n = int(K.sys_n[sid])
                m = n + K.n_border
                coef_i = torch.zeros(_pad_up(m, _lib.SPX_BM) * K.kpad, dtype=_F64,
                                     device=self.device)
                self._lu_solve(ctx, K, T, np.full(m, k), np.full(m, 2), np.arange(m),
                               np.arange(m), coef_i)
                grp = int(K.sys_grp[sid])
                kidx = np.full(_pad_up(m, _lib.SPX_BM), -1, dtype=np.int32)
                kidx[:n] = np.where(ctx['grp_mask'][grp])[0]
                kidx[n:m] = n_stn + np.arange(K.n_border)
                g = self._gemm(ctx, coef=coef_i, n_rows=m, kpad=K.kpad, n_border=K.n_border,
                               gen=_lib.GEN_VG, epi=_lib.EPI_QUADFORM, row_dst=self._dev(kidx),
                               aux=aux, vg=_lib.make_vg(K.uniq_vgs[int(K.sys_vg[sid])]),
                               covar_flag=0, cell_drift=K.d_cell_drift, quad_slot=slot)
            sel = np.where((K.sys_o >= ids[0]) & (K.sys_o <= ids[-1]))[0]
            d_slot = self._dev((K.sys_o[sel] - ids[0]).astype(np.int32))
            d_dst = self._dev(K.steps_o[sel].astype(np.int32))
            _lib.check(self.lib.spx_bcast_rows_dev(
                self._ptr(aux), self._ptr(d_slot), self._ptr(d_dst), int(sel.size), None, None,
                n_cells, self._ptr(ctx['d_pos']), self._ptr(ev_out), ctx['fld_size'],
                ctx['out_f64'], self._stream()), 'bcast_rows')
            self._count('launches')

    # ---- direct path: assemble + LU + substitution per system --------------
    def _systems_struct(self, ctx, K, sys_ids, grp_ids, vg_ids):
        """Device descriptors for a batch of systems (groups may be K.full_grp)."""
        n_stn = ctx['n_stn']
        nb = len(sys_ids)
        grp_n_ext = np.concatenate([ctx['grp_n'], [n_stn]])
        n = grp_n_ext[grp_ids].astype(np.int32)
        m = n.astype(np.int64) + K.n_border
        w_off = np.concatenate([[0], np.cumsum(m * m)])[:-1].astype(np.int64)
        p_off = np.concatenate([[0], np.cumsum(m)])[:-1].astype(np.int64)
        T = types.SimpleNamespace()
        T.work = torch.empty(int((m * m).sum()), dtype=_F64, device=self.device)
        T.piv = torch.empty(int(m.sum()), dtype=_I32, device=self.device)
        T.info = torch.zeros(nb, dtype=_I32, device=self.device)
        T.t = [self._dev(n), self._dev(np.full(nb, K.kind, dtype=np.int32)),
               self._dev(np.asarray(vg_ids, dtype=np.int32)), self._dev(K.stn_off[grp_ids]),
               self._dev(w_off), self._dev(p_off)]
        S = _lib.spx_systems()
        S.n_sys = nb
        S.n_drifts = K.n_drifts
        (S.sys_n, S.sys_kind, S.sys_vg, S.sys_stn_off, S.sys_w_off, S.sys_piv_off) = (
            x.data_ptr() for x in T.t)
        S.stn_list = K.d_stn_list.data_ptr()
        S.stn_x = ctx['d_stn_x'].data_ptr()
        S.stn_y = ctx['d_stn_y'].data_ptr()
        S.stn_drift = K.d_stn_drift.data_ptr() if K.d_stn_drift is not None else None
        S.work = T.work.data_ptr()
        S.piv = T.piv.data_ptr()
        S.info = T.info.data_ptr()
        S.max_m = int(m.max())
        T.S = S
        T.m = m
        return T

    def _factor(self, ctx, K, T):
        _lib.check(self.lib.spx_krige_assemble_dev(
            C.byref(T.S), self._ptr(K.d_vgs), len(K.uniq_vgs), ctx['min_vg_val'], self._stream()),
            'assemble')
        _lib.check(self.lib.spx_krige_factor_dev(C.byref(T.S), self._stream()), 'factor')
        self._count('launches', 2)
        self._count('lu_flop', int((2 * T.m ** 3 // 3).sum()))

    def _lu_solve(self, ctx, K, T, rhs_sys, rhs_kind, rhs_arg, rhs_row, coef, want_resid=False,
                  dense=None, dense_ld=0, row_major=False):
        n_rhs = len(rhs_sys)
        resid = torch.zeros(n_rhs, dtype=_F64, device=self.device) if want_resid else None
        ts = [self._dev(np.asarray(rhs_sys, dtype=np.int32)),
              self._dev(np.asarray(rhs_kind, dtype=np.int32)),
              self._dev(np.asarray(rhs_arg, dtype=np.int32)),
              self._dev(np.asarray(rhs_row, dtype=np.int64))]
        R = _lib.spx_rhs()
        R.n_rhs = n_rhs
        R.rhs_sys, R.rhs_kind, R.rhs_arg, R.rhs_row = (x.data_ptr() for x in ts)
        R.data = ctx['d_data'].data_ptr()
        R.n_stn = ctx['n_stn']
        R.kpad = K.kpad
        R.coef = coef.data_ptr()
        R.resid = resid.data_ptr() if resid is not None else None
        R.dense = dense.data_ptr() if dense is not None else None
        R.dense_ld = int(dense_ld)
        R.coef_row_major = int(bool(row_major))
        _lib.check(self.lib.spx_krige_solve_dev(C.byref(T.S), C.byref(R), self._stream()), 'solve')
        self._count('launches')
        return resid

    def _solve_direct(self, ctx, K, sys_ids, want_resid=True):
        """Assemble, factor and solve the listed systems in batches bounded by the
        workspace limit; data rows go to K.coef, one ones-vector per system gives
        the sum(lambda) residual.  Workspaces are kept in K.keep for the flagged
        handler."""
        m_all = K.sys_n[sys_ids].astype(np.int64) + K.n_border
        nbytes = m_all * m_all * 8
        finishers = []
        b0 = 0
        while b0 < sys_ids.size:
            b1 = b0 + 1
            tot = int(nbytes[b0])
            while (b1 < sys_ids.size and tot + int(nbytes[b1]) <= self.work_limit
                   and (b1 - b0) < 60000):
                tot += int(nbytes[b1])
                b1 += 1
            ids = sys_ids[b0:b1]
            nb = ids.size
            T = self._systems_struct(ctx, K, ids, K.sys_grp[ids], K.sys_vg[ids])
            self._factor(ctx, K, T)
            ridx = K.rows_by_sys[np.isin(K.sys_o[K.rows_by_sys], ids)]   # grouped by system
            local = np.full(K.sys_n.size, -1, dtype=np.int64)
            local[ids] = np.arange(nb)
            n_data = ridx.size
            rhs_sys = np.concatenate([local[K.sys_o[ridx]], np.arange(nb)])
            rhs_kind = np.concatenate([np.zeros(n_data), np.ones(nb)])
            rhs_arg = np.concatenate([K.steps_o[ridx], np.zeros(nb)])
            rhs_row = np.concatenate([K.row_of[ridx], np.full(nb, -1)])
            if not want_resid:
                rhs_sys, rhs_kind = rhs_sys[:n_data], rhs_kind[:n_data]
                rhs_arg, rhs_row = rhs_arg[:n_data], rhs_row[:n_data]
            resid = None
            if rhs_sys.size:
                resid = self._lu_solve(ctx, K, T, rhs_sys, rhs_kind, rhs_arg, rhs_row, K.coef,
                                       want_resid=want_resid, row_major=K.row_major)
            for k, sid in enumerate(ids):
                K.keep[int(sid)] = (T, k)

            h_info = self._fetch_async(T.info)
            h_resid = self._fetch_async(resid[n_data:]) if want_resid else None

            def finish(resid_out, singular_out, ids=ids, h_info=h_info, h_resid=h_resid):
                singular_out[ids] = h_info.numpy() != 0
                if h_resid is not None:
                    resid_out[ids] = h_resid.numpy()
                return False
            finishers.append(finish)
            b0 = b1
        return finishers

    # ---- downdated path -------------------------------------------------
    def _solve_downdate(self, ctx, K, sys_ids):
        """A_g^-1 b from the inverse of the full system of each variogram
        (include/spx_b200.h: spx_downdate).  Only queues work; returns finisher
        callables that later read the health flags (True = something was unhealthy
        and the caller must redo the solve without downdating)."""
        lib = self.lib
        n_stn = ctx['n_stn']
        M = n_stn + K.n_border
        finishers = []
        vgs_here = np.unique(K.sys_vg[sys_ids])
        # Inverse of the full system of every variogram.  It depends on the station
        # set and the variogram only (not on the data), so it is kept across the
        # chunks of a job in a small cache.
        base_key = self._ginv_key(ctx, K)
        ginv_of = {}
        need = []
        for v in vgs_here:
            hit = self._ginv_cache.get((base_key, K.uniq_vgs[int(v)])) if self.ginv_cache else None
            if hit is not None:
                ginv_of[int(v)] = hit
                self._count('ginv_cache_hits')
            else:
                need.append(int(v))
        if need:
            need = np.asarray(need)
            T = self._systems_struct(ctx, K, np.arange(need.size),
                                     np.full(need.size, K.full_grp), need)
            self._factor(ctx, K, T)
            n_full = need.size
            rhs_sys = np.repeat(np.arange(n_full), M + 1)
            rhs_kind = np.tile(np.concatenate([np.full(M, 2), [1]]), n_full)
            rhs_arg = np.tile(np.concatenate([np.arange(M), [0]]), n_full)
            rhs_row = np.full(rhs_sys.size, -1)
            dense = torch.empty((n_full * (M + 1), M), dtype=_F64, device=self.device)
            resid = self._lu_solve(ctx, K, T, rhs_sys, rhs_kind, rhs_arg, rhs_row, K.coef,
                                   want_resid=True, dense=dense, dense_ld=M)
            h_full_info = self._fetch_async(T.info)
            h_full_resid = self._fetch_async(resid.view(n_full, M + 1)[:, M].contiguous())
            dense = dense.view(n_full, M + 1, M)
            for vi, v in enumerate(need):
                ginv_of[int(v)] = dense[vi, :M, :]

            def finish_full(resid_out, singular_out):
                bad = (h_full_info.numpy() != 0) | ~(h_full_resid.numpy() <= 1e-9)
                if self.ginv_cache:
                    for vi, v in enumerate(need):
                        if not bad[vi]:
                            while len(self._ginv_cache) >= self.ginv_cache_size:
                                self._ginv_cache.pop(next(iter(self._ginv_cache)))
                            self._ginv_cache[(base_key, K.uniq_vgs[int(v)])] = ginv_of[int(v)]
                return bool(bad.any())
            finishers.append(finish_full)
        for vi, v in enumerate(vgs_here):
            ids = sys_ids[K.sys_vg[sys_ids] == v]
            if not ids.size:
                continue
            G = ginv_of[int(v)]
            grp = K.sys_grp[ids]
            r = (n_stn - K.sys_n[ids]).astype(np.int32)
            miss_off = np.concatenate([[0], np.cumsum(r)])[:-1].astype(np.int64)
            d_miss_list = self._mask_lists(ctx, grp.astype(np.int32), miss_off, int(r.sum()),
                                           want=0)
            # right-hand sides: data rows of every system + one ones-vector each
            ridx = K.rows_by_sys[np.isin(K.sys_o[K.rows_by_sys], ids)]   # grouped by system
            cnt = (K.sys_row_beg[ids + 1] - K.sys_row_beg[ids]).astype(np.int64)
            n_data = ridx.size
            nsys = ids.size
            # Bt rows: [data rows in ridx order | group masks]
            d_steps = self._dev(K.steps_o[ridx].astype(np.int64))
            self._sync_uploads()
            Bt = torch.zeros((n_data + nsys, M), dtype=_F64, device=self.device)
            Bt[:n_data, :n_stn] = ctx['d_data0'].index_select(0, d_steps)
            d_grp = self._dev(grp.astype(np.int64))
            self._sync_uploads()
            d_maskf = ctx['d_grp_mask'].index_select(0, d_grp)
            Bt[n_data:, :n_stn] = d_maskf.to(_F64)
            Ut = torch.matmul(Bt, G)
            self._count('launches')
            # per-system contiguous rhs lists: its data rows then its ones-vector
            beg = np.concatenate([[0], np.cumsum(cnt)])[:-1]
            rhs_off = (beg + np.arange(nsys)).astype(np.int64)
            rhs_cnt = (cnt + 1).astype(np.int32)
            n_rhs = int(n_data + nsys)
            urow = np.empty(n_rhs, dtype=np.int32)
            rrow = np.empty(n_rhs, dtype=np.int64)
            rkind = np.zeros(n_rhs, dtype=np.int32)
            pos_data = np.arange(n_data) + np.repeat(np.arange(nsys), cnt)
            urow[pos_data] = np.arange(n_data)
            rrow[pos_data] = K.row_of[ridx]
            pos_ones = rhs_off + cnt
            urow[pos_ones] = n_data + np.arange(nsys)
            rrow[pos_ones] = -1
            rkind[pos_ones] = 1
            d_resid = torch.zeros(n_rhs, dtype=_F64, device=self.device)
            d_info = torch.zeros(nsys, dtype=_I32, device=self.device)
            order = np.argsort(-r, kind='stable').astype(np.int32)   # largest systems first
            ts = self._dev_pack([r, miss_off, np.zeros(1, dtype=np.int32), K.sys_n[ids],
                                 K.stn_off[grp], rhs_off, rhs_cnt, urow, rrow, rkind, order])
            d_order = ts.pop()
            ts[2] = d_miss_list
            D = _lib.spx_downdate()
            D.n_sys = nsys
            D.n_stn = n_stn
            D.n_border = K.n_border
            D.max_r = int(r.max()) if r.size else 0
            D.ginv = G.data_ptr()
            (D.sys_r, D.sys_miss_off, D.miss_list, D.sys_n, D.sys_stn_off, D.sys_rhs_off,
             D.sys_rhs_cnt, D.rhs_urow, D.rhs_row, D.rhs_kind) = (x.data_ptr() for x in ts)
            D.stn_list = K.d_stn_list.data_ptr()
            D.ut = Ut.data_ptr()
            D.kpad = K.kpad
            D.coef = K.coef.data_ptr()
            D.coef_row_major = int(K.row_major)
            D.sys_order = d_order.data_ptr()
            D.resid = d_resid.data_ptr()
            D.info = d_info.data_ptr()
            _lib.check(lib.spx_krige_downdate_dev(C.byref(D), self._stream()), 'downdate')
            self._count('launches')

            h_info = self._fetch_async(d_info)
            h_resid = self._fetch_async(d_resid)

            def finish(resid_out, singular_out, ids=ids, h_info=h_info, h_resid=h_resid,
                       pos_ones=pos_ones):
                resid_out[ids] = h_resid.numpy()[pos_ones]
                return bool((h_info.numpy() != 0).any())
            finishers.append(finish)
        return finishers

    def _krige_flagged(self, ctx, out, K, flagged, singular, bad_cells, problem_steps):
        """Systems whose computed weights may not sum to one (always for SK,
        quirk Q6): evaluate sum(lambda) per cell exactly like steps.py:418 and
        overwrite failing cells with the nearest available station."""
        lib = self.lib
        n_cells = ctx['n_cells']
        fl = np.where(flagged)[0]
        max_slots = max(1, int(self.aux_limit // (n_cells * 13)))
        d_cell_bad = None
        if bad_cells.size:
            cb = np.zeros(n_cells, dtype=np.uint8)
            cb[bad_cells] = 1
            d_cell_bad = self._dev(cb)
        for c0 in range(0, fl.size, max_slots):
            fb = fl[c0:c0 + max_slots]
            n_f = fb.size
            fail = torch.ones((n_f, n_cells), dtype=torch.uint8, device=self.device)
            aux = torch.empty((n_f, n_cells), dtype=_F64, device=self.device)
            # OK / EDK systems end up here only when LU cannot be trusted (zero pivot or
            # a large ones-vector residual): redo them with the reference's own
            # operator, the SVD pseudo-inverse with rcond = 1e-15 (np.linalg.pinv,
            # steps.py:351) -- for the symmetric A it is V diag(1/w) V' over the
            # eigenvalues with |w| > rcond * max|w| -- so that truncated systems give
            # the reference's weights, its sum(lambda) verdict and its estimates.
            use_pinv = np.zeros(flagged.size, dtype=bool)
            if K.kind != 1 and self.pinv_flagged:
                use_pinv[fb] = True
            ok = fb[~singular[fb] | use_pinv[fb]]
            for v in np.unique(K.sys_vg[ok]):
                sv = ok[K.sys_vg[ok] == v]
                # compact variogram: the ones-vector rows go through the local estimator
                # (stations within range + the constant far field) instead of a contraction
                # that regenerates the whole cells x stations tile for a handful of rows
                loc = None
                if (getattr(K, 'local', None) is not None and K.n_drifts == 0
                        and self.local_support):
                    loc = self._local_plan(ctx, [K.uniq_vgs[int(v)]])
                rows_a = int(sv.size) if loc is not None else _pad_up(sv.size, _lib.SPX_BM)
                coef_a = torch.zeros(rows_a * K.kpad, dtype=_F64, device=self.device)
                sv_lu = sv[~use_pinv[sv]]
                sv_pi = sv[use_pinv[sv]]
                # ones-vector solutions, grouped by the kept factor batch
                by_T = {}
                for s in sv_lu:
                    row = int(np.searchsorted(sv, s))
                    T, k = K.keep[int(s)]
                    by_T.setdefault(id(T), (T, [], []))
                    by_T[id(T)][1].append(k)
                    by_T[id(T)][2].append(row)
                for T, ks, rows in by_T.values():
                    self._lu_solve(ctx, K, T, ks, np.ones(len(ks)), np.zeros(len(ks)), rows,
                                   coef_a, row_major=loc is not None)
                if sv_pi.size:
                    self._pinv_systems(ctx, out, K, sv_pi, sv, int(v), coef_a,
                                       row_major=loc is not None)
                d_slots = self._dev(np.searchsorted(fb, sv).astype(np.int32))
                if loc is not None:
                    self._local_aux(ctx, K, K.uniq_vgs[int(v)], loc[0], coef_a, int(sv.size),
                                    d_slots, aux)
                    continue
                self._gemm(ctx, coef=coef_a, n_rows=int(sv.size), kpad=K.kpad,
                           n_border=K.n_border, gen=_lib.GEN_VG, epi=_lib.EPI_AUX,
                           row_dst=d_slots, aux=aux, vg=_lib.make_vg(K.uniq_vgs[int(v)]),
                           covar_flag=int(K.kind == 1), cell_drift=K.d_cell_drift)
            if ok.size:
                ok_slots = self._dev(np.searchsorted(fb, ok).astype(np.int64))
                self._sync_uploads()
                aux_ok = aux[ok_slots]
                fail_ok = torch.empty((ok.size, n_cells), dtype=torch.uint8, device=self.device)
                _lib.check(lib.spx_lambda_check_dev(
                    self._ptr(aux_ok), int(ok.size), n_cells, self._ptr(d_cell_bad),
                    self._ptr(fail_ok), self._stream()), 'lambda_check')
                self._count('launches')
                fail[ok_slots] = fail_ok
            slot_of_sys = np.full(flagged.size, -1, dtype=np.int64)
            slot_of_sys[fb] = np.arange(n_f)
            rsel = np.where(slot_of_sys[K.sys_o] >= 0)[0]
            if not rsel.size:
                continue
            r_steps = K.steps_o[rsel]
            gl = np.unique(K.sys_grp[fb])
            gslot = {int(g): i for i, g in enumerate(gl)}
            nnb = self._nnb_index(ctx, gl)
            self._nnb_gather(ctx, out, nnb, r_steps,
                             [gslot[int(ctx['grp_of_step'][s])] for s in r_steps],
                             fail=fail, row_fail=slot_of_sys[K.sys_o[rsel]])
            if K.ev_out is not None:                       # steps.py:425-426, :384
                d_rs = self._dev(r_steps.astype(np.int32))
                d_rf = self._dev(slot_of_sys[K.sys_o[rsel]].astype(np.int32))
                _lib.check(lib.spx_bcast_rows_dev(
                    None, None, self._ptr(d_rs), int(r_steps.size), self._ptr(fail),
                    self._ptr(d_rf), n_cells, self._ptr(ctx['d_pos']), self._ptr(K.ev_out),
                    ctx['fld_size'], ctx['out_f64'], self._stream()), 'bcast_rows(zero)')
                self._count('launches')
            for s in fb[singular[fb] & ~use_pinv[fb]]:
                for t in K.steps_o[K.sys_o == s]:
                    if int(t) not in problem_steps:
                        problem_steps.append(int(t))
