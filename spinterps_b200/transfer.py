"""Device -> host transport of rounded float32 fields in two bytes per value.

The writer stores ``np.round(fld, nmrl_prcn)`` as float32 (reference
interp/steps.py:907-945).  Such a field is, row by row (time step by time step), an
integer lattice ``q / 10**d``; ``spx_pack_field_dev`` (csrc/spx_pack.cu) turns every row
that round-trips bit for bit into 16-bit codes, the copy over PCIe carries half the bytes,
and ``spx_unpack_field_host`` rebuilds the identical floats on the host cores (threaded,
AVX2).  Rows that do not qualify stay float32 and are copied as they are.  Nothing is
approximated: ``finish()`` returns exactly the bytes ``fld.cpu()`` would.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class PackedDownloader:
    """Ring of ``depth`` transfer slots for fields of at most ``max_rows`` x ``row_len``.

    ``start(fld, decimals)`` queues the pack kernels on the current stream and the copies
    on the downloader's copy stream and returns a ticket; ``finish(ticket, out)`` waits
    for the copies, fetches raw rows and decodes into ``out`` (float32 [rows, row_len],
    any row pitch).  Several tickets may be in flight (at most ``depth``)."""

    def __init__(self, device, max_rows, row_len, depth=2, n_threads=0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.max_rows, self.row_len = int(max_rows), int(row_len)
        self.stride = int(self.lib.spx_pack_stride(self.row_len))
        self.n_threads = int(n_threads)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = []
        for _ in range(int(depth)):
            self.slots.append(dict(
                d_hdr=torch.empty(self.max_rows * 4, dtype=torch.int32, device=self.device),
                d_codes=torch.empty(self.max_rows * self.stride, dtype=torch.int16,
                                    device=self.device),
                h_hdr=torch.empty(self.max_rows * 4, dtype=torch.int32).pin_memory(),
                h_codes=torch.empty(self.max_rows * self.stride, dtype=torch.int16).pin_memory(),
                ev=torch.cuda.Event(), fld=None, busy=False))
        self._next = 0
        self.d2h_bytes = 0          # bytes copied device -> host so far

    def bytes_per_field(self, n_rows):
        """Bytes the packed copy of an all-qualifying field moves (header + codes)."""
        return int(n_rows) * (16 + 2 * self.stride)

    def start(self, fld, decimals):
        assert fld.dtype == torch.float32 and fld.dim() == 2 and fld.stride(1) == 1
        n_rows, row_len = fld.shape
        assert n_rows <= self.max_rows and row_len == self.row_len
        k = self._next
        s = self.slots[k]
        assert not s['busy'], 'more tickets in flight than slots'
        self._next = (k + 1) % len(self.slots)
        main = torch.cuda.current_stream(self.device)
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

    def wait(self, ticket):
        """Wait for the ticket's copies and return the field in its packed form
        (PackedField over the slot's pinned buffers: valid until ``release(ticket)`` or the
        next ``start`` that reuses the slot).  Rows that travelled as raw floats are fetched
        here."""
        s = self.slots[ticket]
        assert s['busy']
        n_rows = s['n_rows']
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
        self.slots[ticket].update(fld=None, busy=False)

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
