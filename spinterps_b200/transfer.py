"""Device -> host transport of rounded float32 fields in a lossless compact form.

The writer stores ``np.round(fld, nmrl_prcn)`` as float32 (reference
interp/steps.py:907-945).  Such a field is an integer lattice ``q / 10**d``.  Two codecs
(csrc/spx_pack.cu), both verified bit for bit on the device, value by value:

* ``'delta'`` (default): per tile of 256 cells the chain of q is delta coded and bit-packed
  with one width per 8 cells (``spx_dpack_field_dev``).  An interpolated field is smooth, so
  it needs a few bits per cell: 0.3-0.6 bytes instead of 4.  Variable size: the records of
  a field land in one payload buffer, the copy over PCIe is issued for the bytes actually
  used as soon as the encoder's byte count is on the host.
* ``'u16'``: 16-bit codes per row (``spx_pack_field_dev``), 2 bytes per cell; also the
  fallback when the delta form of a field does not fit its buffer (noise-like fields).

``spx_dunpack_rows_host`` / ``spx_unpack_field_host`` rebuild the identical floats on the
host cores (threaded).  Nothing is approximated: ``finish()`` returns exactly the bytes
``fld.cpu()`` would.
"""
from __future__ import annotations

import concurrent.futures
import ctypes as C
import os

import numpy as np
import torch

from . import _lib


def default_codec():
    c = os.environ.get('SPX_TRANSPORT', 'delta')
    if c not in ('delta', 'u16'):
        raise ValueError("SPX_TRANSPORT must be 'delta' or 'u16'")
    return c


class PackedDownloader:
    """Ring of ``depth`` transfer slots for fields of at most ``max_rows`` x ``row_len``.

    ``start(fld, decimals)`` queues the encoder on the current stream and the copies on the
    downloader's copy streams and returns a ticket; ``wait(ticket)`` returns the field in its
    packed host form, ``finish(ticket, out)`` also decodes it into ``out`` (float32
    [rows, row_len], any row pitch).  Several tickets may be in flight (at most ``depth``).

    codec 'delta': ``bytes_per_cell`` sizes the payload buffers (device and pinned host);
    a field whose records need more is sent through the 16-bit codec instead."""

    def __init__(self, device, max_rows, row_len, depth=2, n_threads=0, codec=None,
                 bytes_per_cell=1.25):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.max_rows, self.row_len = int(max_rows), int(row_len)
        self.depth = int(depth)
        self.codec = codec or default_codec()
        self.n_threads = int(n_threads)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.stride = int(self.lib.spx_pack_stride(self.row_len))
        self.segs = int(self.lib.spx_dpack_segments(self.row_len))
        self.slots = []
        self._u16 = None                 # fallback downloader (allocated on first use)
        self.fallbacks = 0               # fields that did not fit their delta buffer
        if self.codec == 'delta':
            worst = int(self.lib.spx_dpack_capacity(self.max_rows, self.row_len))
            cap = int(bytes_per_cell * self.max_rows * self.row_len)
            self.capacity = max(4096, min(worst, (cap + 4095) // 4096 * 4096))
            self.pay_stream = torch.cuda.Stream(self.device)
            self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1)
            n_off = self.max_rows * self.segs
            for _ in range(self.depth):
                self.slots.append(dict(
                    d_off=torch.empty(n_off, dtype=torch.int32, device=self.device),
                    d_pay=torch.empty(self.capacity, dtype=torch.uint8, device=self.device),
                    d_cnt=torch.zeros(2, dtype=torch.int64, device=self.device),
                    h_off=torch.empty(n_off, dtype=torch.int32).pin_memory(),
                    h_pay=torch.empty(self.capacity, dtype=torch.uint8).pin_memory(),
                    h_cnt=torch.zeros(2, dtype=torch.int64).pin_memory(),
                    ev_meta=torch.cuda.Event(), ev=torch.cuda.Event(), fld=None, busy=False,
                    fut=None, overflow=False, pay_bytes=0))
        else:
            for _ in range(self.depth):
                self.slots.append(dict(
                    d_hdr=torch.empty(self.max_rows * 4, dtype=torch.int32, device=self.device),
                    d_codes=torch.empty(self.max_rows * self.stride, dtype=torch.int16,
                                        device=self.device),
                    h_hdr=torch.empty(self.max_rows * 4, dtype=torch.int32).pin_memory(),
                    h_codes=torch.empty(self.max_rows * self.stride,
                                        dtype=torch.int16).pin_memory(),
                    ev=torch.cuda.Event(), fld=None, busy=False))
        self._next = 0
        self.d2h_bytes = 0          # bytes copied device -> host so far

    def bytes_per_field(self, n_rows):
        """Bytes the 16-bit copy of an all-qualifying field moves (header + codes)."""
        return int(n_rows) * (16 + 2 * self.stride)

    # ------------------------------------------------------------------ start
    def start(self, fld, decimals, round_here=False, stats=None, write_back=True):
        """Queue encoder + copies for ``fld`` (device float32 [n_rows, row_len], rounded to
        ``decimals`` places).  codec 'delta' only: ``round_here=True`` takes the UNROUNDED
        field and makes the call the whole output stage of the writer in one pass --
        np.round, the per-step statistics into ``stats`` (device float64 [5, n_rows], or
        None) and the encoding; ``write_back`` also leaves the rounded values in ``fld``."""
        assert fld.dtype == torch.float32 and fld.dim() == 2 and fld.stride(1) == 1
        n_rows, row_len = fld.shape
        assert n_rows <= self.max_rows and row_len == self.row_len
        assert self.codec == 'delta' or (not round_here and stats is None)
        k = self._next
        s = self.slots[k]
        assert not s['busy'], 'more tickets in flight than slots'
        self._next = (k + 1) % len(self.slots)
        main = torch.cuda.current_stream(self.device)
        if self.codec == 'delta':
            flags = (_lib.SPX_DPACK_ROUND if round_here else 0) | (
                _lib.SPX_DPACK_WRITE_BACK if (round_here and write_back) else 0)
            ws = None
            if stats is not None:
                assert stats.dtype == torch.float64 and tuple(stats.shape) == (5, n_rows) \
                    and stats.is_contiguous()
                ws = torch.empty(max(1, int(self.lib.spx_dpack_stats_workspace(n_rows, row_len))),
                                 dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.spx_dpack_field_dev(
                C.c_void_p(fld.data_ptr()), n_rows, row_len, fld.stride(0), int(decimals), flags,
                C.c_void_p(stats.data_ptr() if stats is not None else None),
                C.c_void_p(ws.data_ptr() if ws is not None else None),
                C.c_void_p(s['d_off'].data_ptr()), C.c_void_p(s['d_pay'].data_ptr()),
                self.capacity, C.c_void_p(s['d_cnt'].data_ptr()),
                C.c_void_p(main.cuda_stream)), 'dpack_field')
            s['rounded_in_place'] = bool(round_here and write_back)
            s['round_here'] = bool(round_here)
            self.copy_stream.wait_stream(main)
            n_off = n_rows * self.segs
            with torch.cuda.stream(self.copy_stream):
                s['h_cnt'].copy_(s['d_cnt'], non_blocking=True)
                s['h_off'][:n_off].copy_(s['d_off'][:n_off], non_blocking=True)
                s['ev_meta'].record(self.copy_stream)
            s.update(fld=fld, n_rows=int(n_rows), decimals=int(decimals), busy=True,
                     overflow=False, pay_bytes=0)
            s['fut'] = self._pool.submit(self._land, k)
            return k
        _lib.check(self.lib.spx_pack_field_dev(
            C.c_void_p(fld.data_ptr()), n_rows, row_len, fld.stride(0), int(decimals),
            C.c_void_p(s['d_hdr'].data_ptr()), C.c_void_p(s['d_codes'].data_ptr()),
            C.c_void_p(main.cuda_stream)), 'pack_field')
        self.copy_stream.wait_stream(main)
        with torch.cuda.stream(self.copy_stream):
            s['h_hdr'][:n_rows * 4].copy_(s['d_hdr'][:n_rows * 4], non_blocking=True)
            s['h_codes'][:n_rows * self.stride].copy_(s['d_codes'][:n_rows * self.stride],
                                                      non_blocking=True)
            s['ev'].record(self.copy_stream)
        fld.record_stream(self.copy_stream)
        s.update(fld=fld, n_rows=int(n_rows), decimals=int(decimals), busy=True)
        self.d2h_bytes += n_rows * (16 + 2 * self.stride)
        return k

    def _land(self, k):
        """Helper thread: as soon as the encoder's byte count is on the host, queue the copy
        of exactly the bytes it used."""
        s = self.slots[k]
        torch.cuda.set_device(self.device)
        s['ev_meta'].synchronize()
        nbytes = int(s['h_cnt'][0]) * 4
        meta = 16 + s['n_rows'] * self.segs * 4
        if int(s['h_cnt'][1]) != 0 or nbytes > self.capacity:
            s['overflow'] = True
            self.d2h_bytes += meta
            return
        if nbytes:
            with torch.cuda.stream(self.pay_stream):
                s['h_pay'][:nbytes].copy_(s['d_pay'][:nbytes], non_blocking=True)
        s['ev'].record(self.pay_stream)
        s['pay_bytes'] = nbytes
        self.d2h_bytes += meta + nbytes

    # ------------------------------------------------------------------ wait
    def wait(self, ticket):
        """Wait for the ticket's copies and return the field in its packed form (DeltaField
        / PackedField over the slot's pinned buffers: valid until ``release(ticket)`` or the
        next ``start`` that reuses the slot)."""
        s = self.slots[ticket]
        assert s['busy']
        n_rows = s['n_rows']
        if self.codec == 'delta':
            s['fut'].result()
            if s['overflow']:
                # noise-like field: the 16-bit codec (or raw rows) carries it
                self.fallbacks += 1
                if self._u16 is None:
                    self._u16 = PackedDownloader(self.device, self.max_rows, self.row_len,
                                                 depth=1, n_threads=self.n_threads, codec='u16')
                u = self._u16
                d0 = u.d2h_bytes
                fld = s['fld']
                if s['round_here'] and not s['rounded_in_place']:
                    # the fused pass left the field unrounded: round it now (np.round)
                    ws = torch.empty(max(1, int(self.lib.spx_round_stats_workspace(
                        fld.shape[0], fld.shape[1]))), dtype=torch.uint8, device=self.device)
                    tmp = torch.empty((5, fld.shape[0]), dtype=torch.float64, device=self.device)
                    _lib.check(self.lib.spx_round_stats_dev(
                        C.c_void_p(fld.data_ptr()), 0, fld.shape[0], fld.shape[1], fld.stride(0),
                        s['decimals'], C.c_void_p(tmp.data_ptr()), C.c_void_p(ws.data_ptr()),
                        C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                        'round_stats')
                pf = u.wait(u.start(fld, s['decimals']))
                self.d2h_bytes += u.d2h_bytes - d0
                s['fld'] = None
                return pf
            s['ev'].synchronize()
            s['fld'] = None
            offs = s['h_off'].numpy()[:n_rows * self.segs].view(np.uint32)
            pay = s['h_pay'].numpy()[:s['pay_bytes']]
            return DeltaField(self.lib, offs, pay, n_rows, self.row_len, s['decimals'])
        s['ev'].synchronize()
        hdr = s['h_hdr'].numpy()[:n_rows * 4].view(_lib.PACK_ROW_DTYPE)
        codes = s['h_codes'].numpy()[:n_rows * self.stride].view(np.uint16).reshape(
            n_rows, self.stride)
        raw_rows = np.where(hdr['mode'] != _lib.SPX_PACK_U16)[0]
        raw = {}
        if raw_rows.size:
            fld = s['fld']
            with torch.cuda.stream(self.copy_stream):
                for r in raw_rows:
                    raw[int(r)] = fld[int(r)].cpu().numpy()
            self.d2h_bytes += int(raw_rows.size) * self.row_len * 4
        s['fld'] = None
        return PackedField(self.lib, hdr, codes, raw, self.row_len, s['decimals'])

    def release(self, ticket):
        s = self.slots[ticket]
        if self._u16 is not None and self._u16.slots[0]['busy']:
            self._u16.release(0)
        s.update(fld=None, busy=False)

    def finish(self, ticket, out):
        """Decode the ticket's field into ``out`` (ndarray float32 [n_rows, row_len]);
        returns the number of rows that travelled as raw floats."""
        pf = self.wait(ticket)
        pf.decode(out, self.n_threads)
        self.release(ticket)
        return len(pf.raw)

    def download(self, fld, decimals, out=None):
        """start + finish; returns the host array."""
        if out is None:
            out = np.empty(tuple(fld.shape), dtype=np.float32)
        self.finish(self.start(fld, decimals), out)
        return out


class PackedField:
    """A rounded float32 field [n_rows, row_len] in host memory in its lossless 2-byte
    form: per row a header (mode, qmin) and 16-bit codes, raw float rows where a row did
    not qualify.  ``row(i)`` / ``decode()`` rebuild exactly the floats the device held; a
    writer decodes one (1, ny, nx) step at a time, in cache, right before it compresses
    it, so the float field never has to exist in host memory as a whole."""

    codec = 'u16'

    def __init__(self, lib, hdr, codes, raw, row_len, decimals):
        self.lib, self.hdr, self.codes, self.raw = lib, hdr, codes, raw
        self.row_len, self.decimals = int(row_len), int(decimals)
        self.shape = (int(hdr.shape[0]), self.row_len)
        self.dtype = np.dtype(np.float32)
        self.release = lambda: None          # set by the producer: frees the transfer slot
        self.nbytes = int(hdr.nbytes + codes.nbytes + sum(v.nbytes for v in raw.values()))

    def decode(self, out=None, n_threads=0):
        n_rows = self.shape[0]
        if out is None:
            out = np.empty(self.shape, dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == self.shape
        assert out.strides[1] == 4 and out.strides[0] % 4 == 0
        _lib.check(self.lib.spx_unpack_field_host(
            self.hdr.ctypes.data, self.codes.ctypes.data, n_rows, self.row_len, self.decimals,
            out.ctypes.data, out.strides[0] // 4, int(n_threads)), 'unpack_field')
        for r, v in self.raw.items():
            out[r] = v
        return out

    def row(self, i, out=None):
        """One row (time step) as float32 [row_len]."""
        i = int(i)
        if i in self.raw:
            if out is None:
                return self.raw[i].copy()
            out[:] = self.raw[i]
            return out
        if out is None:
            out = np.empty(self.row_len, dtype=np.float32)
        _lib.check(self.lib.spx_unpack_field_host(
            self.hdr[i:i + 1].ctypes.data, self.codes[i:i + 1].ctypes.data, 1, self.row_len,
            self.decimals, out.ctypes.data, self.row_len, 1), 'unpack_field')
        return out


class DeltaField:
    """A rounded float32 field [n_rows, row_len] in host memory in the delta transport form
    (tile offsets + one payload of variable-size tile records, csrc/spx_pack.cu).  Same
    interface as PackedField: ``row(i)`` decodes one time step on the calling thread (the
    writer's compression workers call it concurrently), ``decode()`` the whole field."""

    codec = 'delta'

    def __init__(self, lib, offs, payload, n_rows, row_len, decimals):
        self.lib, self.offs, self.payload = lib, offs, payload
        self.row_len, self.decimals = int(row_len), int(decimals)
        self.segs = int(lib.spx_dpack_segments(self.row_len))
        self.shape = (int(n_rows), self.row_len)
        self.dtype = np.dtype(np.float32)
        self.raw = {}                        # no whole row travels raw in this form
        self.release = lambda: None
        self.nbytes = int(offs.nbytes + payload.nbytes)

    def _decode(self, r0, n, out_ptr, out_ld, n_threads):
        _lib.check(self.lib.spx_dunpack_rows_host(
            self.offs[r0 * self.segs:].ctypes.data, self.payload.ctypes.data,
            self.payload.nbytes, n, self.row_len, self.decimals, out_ptr, out_ld,
            int(n_threads)), 'dunpack_rows')

    def decode(self, out=None, n_threads=0):
        if out is None:
            out = np.empty(self.shape, dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == self.shape
        assert out.strides[1] == 4 and out.strides[0] % 4 == 0
        if self.shape[0] and self.row_len:
            self._decode(0, self.shape[0], out.ctypes.data, out.strides[0] // 4, n_threads)
        return out

    def row(self, i, out=None):
        """One row (time step) as float32 [row_len]."""
        if out is None:
            out = np.empty(self.row_len, dtype=np.float32)
        self._decode(int(i), 1, out.ctypes.data, self.row_len, 1)
        return out
