"""Output file with the reference's netCDF layout (interp/prepare.py:290-432).

Dimensions ``dimx, dimy, dimt``; variables ``X`` ('d'), ``Y`` ('d', descending),
``time`` ('i8', + units / calendar); one ``(dimt, dimy, dimx)`` variable of the
field dtype per interpolation label, zlib-compressed (+ shuffle) in ``(1, ny, nx)``
chunks, with ``units`` / ``standard_name``; the 29 ``sett_*`` global attributes and
``Source``.

Two backends write the SAME layout: the ``netCDF4`` package when it is installed, else the
self-contained NetCDF-4 / HDF5 writer of ``nc4file.py`` (the build image has no netCDF4 /
HDF5 library), whose field chunks are compressed by a pool of threads.  Which backend wrote
a file is recorded in the global attribute ``spx_backend``.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import nc4file

try:  # pragma: no cover - not installed in the build image
    import netCDF4 as _nc4
except Exception:  # noqa: BLE001
    _nc4 = None


def have_netcdf4():
    return _nc4 is not None


class _NC4Handle:
    def __init__(self, path, mode):
        self._h = _nc4.Dataset(str(path), mode=mode)
        self._h.set_auto_mask(False)

    def write(self, label, t_idx, row_beg, row_end, values):
        self._h[label][t_idx, row_beg:row_end, :] = values

    def read(self, label, t_idx):
        return np.asarray(self._h[label][t_idx])

    def sync(self):
        self._h.sync()

    def close(self):
        self._h.close()


class _SpxHandle:
    """nc4file.Nc4Writer behind the same four calls.  A file stays open between
    ``open_for_update`` calls of one process (``close`` only syncs: the chunk index is
    written back, the file is complete on disk); ``finalize(path)`` closes it."""

    def __init__(self, path):
        self._w = nc4file.Nc4Writer(path, 'r+')

    def write(self, label, t_idx, row_beg, row_end, values):
        ny = self._w.vars[label]['shape'][1]
        if isinstance(t_idx, slice):
            t0 = t_idx.start or 0
            if row_beg == 0 and row_end == ny:
                self._w.write_steps(label, t0, values)
            else:
                vals = np.asarray(values)
                for i in range(vals.shape[0]):
                    self._w.write_rows(label, t0 + i, row_beg, row_end, vals[i])
        else:
            self._w.write_rows(label, int(t_idx), row_beg, row_end, values)

    def write_packed(self, label, t_index, packed):
        """Whole steps in the 2-byte transport form (transfer.PackedField): row i goes to
        step t_index[i]; decoded step by step inside the compression workers."""
        self._w.write_steps(label, 0, packed, t_index=t_index)

    def read(self, label, t_idx):
        return self._w.read_step(label, int(t_idx))

    def sync(self):
        self._w.sync()

    def close(self):
        self._w.sync()


class _SpxReadHandle:
    def __init__(self, path):
        self._r = nc4file.Nc4Reader(path)

    def read(self, label, t_idx):
        return self._r.read_step(label, int(t_idx))

    def close(self):
        self._r.close()


_OPEN = {}


def open_for_update(path):
    """Handle with write(label, t_idx, row_beg, row_end, values) / sync / close
    (the reference re-opens the file in 'r+' mode per task, steps.py:908)."""
    if _nc4 is not None:
        return _NC4Handle(path, 'r+')
    key = str(Path(path).resolve())
    h = _OPEN.get(key)
    if h is None:
        h = _OPEN[key] = _SpxHandle(path)
    return h


def finalize(path):
    """Close the cached writer of ``path`` (no-op with the netCDF4 backend)."""
    h = _OPEN.pop(str(Path(path).resolve()), None)
    if h is not None:
        h._w.close()


def open_for_read(path):
    if _nc4 is not None:
        return _NC4Handle(path, 'r')
    finalize(path)
    return _SpxReadHandle(path)


def time_numbers(time_rng, units, calendar, tfreq):
    """interp/prepare.py:337-351: date2num, integer-divided by the numeric prefix
    of the frequency string.  Without netCDF4/cftime only the 'X since YYYY-...'
    units with the standard / gregorian calendars are supported."""
    if _nc4 is not None:
        nums = _nc4.date2num(time_rng.to_pydatetime(), units=units, calendar=calendar)
    else:
        import pandas as pd
        unit, _, since = units.partition(' since ')
        assert since, f'unsupported time units: {units!r}'
        assert calendar in ('standard', 'gregorian', 'proleptic_gregorian'), calendar
        delta = (time_rng - pd.Timestamp(since)).to_numpy().astype('timedelta64[s]').astype(
            np.int64)
        div = {'seconds': 1, 'minutes': 60, 'hours': 3600, 'days': 86400}[unit.strip().lower()]
        nums = delta // div
    nums = np.asarray(nums, dtype=np.int64)
    aht_idx = tfreq.find(next(filter(str.isalpha, tfreq)))
    if aht_idx > 0:
        nums //= int(tfreq[:aht_idx])
    return nums


def create(path, x_crds, y_crds, time_vals, interp_args, field_dtype, var_units, var_label,
           time_units=None, time_calendar=None, complevel=1, settings=None,
           xlab='X', ylab='Y', tlab='time'):
    """Create the output file (mode 'w', like interp/prepare.py:308-311) and
    return its path."""
    path = Path(path)
    nx, ny, nt = len(x_crds), len(y_crds), len(time_vals)
    settings = dict(settings or {})
    if _nc4 is not None:
        h = _nc4.Dataset(str(path), mode='w', encoding='utf-8')
        h.set_auto_mask(False)
        h.createDimension('dimx', nx)
        h.createDimension('dimy', ny)
        h.createDimension('dimt', nt)
        h.createVariable(xlab, 'd', dimensions='dimx')[:] = x_crds
        h.createVariable(ylab, 'd', dimensions='dimy')[:] = y_crds
        tv = h.createVariable(tlab, 'i8', dimensions='dimt')
        tv[:] = np.asarray(time_vals, dtype=np.int64)
        if time_units is not None:
            tv.units = time_units
            tv.calendar = time_calendar
        for arg in interp_args:
            name = arg[2]
            v = h.createVariable(name, field_dtype, dimensions=('dimt', 'dimy', 'dimx'),
                                 fill_value=False, compression='zlib', complevel=complevel,
                                 chunksizes=(1, ny, nx))
            v.units = var_units
            v.standard_name = var_label + (
                f' ({name[:3]}_exp_{arg[3]})' if arg[0] == 'IDW' else f' ({name})')
        for k, val in settings.items():
            setattr(h, k, str(val))
        h.spx_backend = 'netCDF4'
        h.Source = str(path)
        h.close()
        return path

    finalize(path)
    variables = [
        dict(name=xlab, dtype='f8', dims=('dimx',), data=np.asarray(x_crds, dtype=np.float64),
             attrs=[]),
        dict(name=ylab, dtype='f8', dims=('dimy',), data=np.asarray(y_crds, dtype=np.float64),
             attrs=[]),
        dict(name=tlab, dtype='i8', dims=('dimt',), data=np.asarray(time_vals, dtype=np.int64),
             attrs=([('units', str(time_units)), ('calendar', str(time_calendar))]
                    if time_units is not None else [])),
    ]
    for arg in interp_args:
        name = arg[2]
        std = var_label + (f' ({name[:3]}_exp_{arg[3]})' if arg[0] == 'IDW' else f' ({name})')
        variables.append(dict(name=name, dtype=np.dtype(field_dtype), dims=('dimt', 'dimy', 'dimx'),
                              data=None, chunk=(1, ny, nx), deflate=int(complevel),
                              attrs=[('units', str(var_units)), ('standard_name', std)]))
    gattrs = [(k, str(val)) for k, val in settings.items()]
    gattrs += [('spx_backend', 'spx-hdf5'), ('Source', str(path))]
    w = nc4file.Nc4Writer(path, 'w')
    w.create([('dimx', nx), ('dimy', ny), ('dimt', nt)], variables, gattrs)
    w.close()
    return path
