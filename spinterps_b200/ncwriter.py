"""Output file with the reference's netCDF layout (interp/prepare.py:290-432).

Dimensions ``dimx, dimy, dimt``; variables ``X`` ('d'), ``Y`` ('d', descending),
``time`` ('i8', + units / calendar); one ``(dimt, dimy, dimx)`` variable of the
field dtype per interpolation label with ``units`` / ``standard_name``; the 29
``sett_*`` global attributes and ``Source``.

With ``netCDF4`` installed the file is NETCDF4 with zlib compression and
``(1, ny, nx)`` chunks exactly like the reference.  ``netCDF4`` is absent in the
build image, so there is a NetCDF-3 (classic, 64-bit offset) fallback through
``scipy.io.netcdf_file`` with the same dimensions, variables and attributes but
no compression, and ``time`` stored as 'i4' pairs is avoided by using 'd' there
(NetCDF-3 has no 64-bit integer).  Which backend wrote a file is recorded in the
global attribute ``spx_backend``.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

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


class _NC3Handle:
    def __init__(self, path, mode):
        from scipy.io import netcdf_file
        self._h = netcdf_file(str(path), mode, mmap=False, version=2)

    def write(self, label, t_idx, row_beg, row_end, values):
        self._h.variables[label][t_idx, row_beg:row_end, :] = values

    def read(self, label, t_idx):
        return np.array(self._h.variables[label][t_idx])

    def sync(self):
        self._h.flush()

    def close(self):
        self._h.close()


def open_for_update(path):
    """Handle with write(label, t_idx, row_beg, row_end, values) / sync / close
    (the reference re-opens the file in 'r+' mode per task, steps.py:908)."""
    if _nc4 is not None:
        return _NC4Handle(path, 'r+')
    return _NC3Handle(path, 'a')


def open_for_read(path):
    if _nc4 is not None:
        return _NC4Handle(path, 'r')
    return _NC3Handle(path, 'r')


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

    from scipy.io import netcdf_file
    h = netcdf_file(str(path), 'w', mmap=False, version=2)
    h.createDimension('dimx', nx)
    h.createDimension('dimy', ny)
    h.createDimension('dimt', nt)
    h.createVariable(xlab, 'd', ('dimx',))[:] = x_crds
    h.createVariable(ylab, 'd', ('dimy',))[:] = y_crds
    tv = h.createVariable(tlab, 'd', ('dimt',))
    tv[:] = np.asarray(time_vals, dtype=np.float64)
    if time_units is not None:
        tv.units = time_units
        tv.calendar = time_calendar
    tcode = 'f' if np.dtype(field_dtype) == np.float32 else 'd'
    for arg in interp_args:
        name = arg[2]
        v = h.createVariable(name, tcode, ('dimt', 'dimy', 'dimx'))
        v[:] = np.nan
        v.units = var_units
        v.standard_name = var_label + (
            f' ({name[:3]}_exp_{arg[3]})' if arg[0] == 'IDW' else f' ({name})')
    for k, val in settings.items():
        setattr(h, k, str(val))
    h.spx_backend = 'scipy-netcdf3'
    h.Source = str(path)
    h.close()
    return path
