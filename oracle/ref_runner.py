"""Run the REFERENCE's own implementation of the hot path, compiled into oracle/_ref/ by
oracle/build_ref.py (Cython + gcc from /root/reference, nothing copied).

TEST INFRASTRUCTURE -- imported only by tests/, __graft_entry__ and bench.py's CPU arm.

``load()`` makes the extension modules importable as package ``spinterps`` (synthetic
package objects instead of the reference's own ``__init__.py`` files, which pull in the
whole GIS stack), with MagicMock stand-ins for the I/O libraries the path never calls
(netCDF4, osgeo, pathos, shapefile, cftime, matplotlib).
``run_case`` builds the fake main object with the 19 attributes read at
interp/steps.py:33-53 and calls ``SpInterpSteps(main)._get_all_interp_outputs(args)``.
"""
import importlib
import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

HERE = Path(__file__).resolve().parent
PKG = HERE / '_ref' / 'spinterps'

_loaded = None


def available():
    from . import build_ref
    return build_ref.available()


def load():
    """-> (interpmthds module, SpInterpSteps, misc module); raises if oracle/_ref is
    missing."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('oracle/_ref is not built (python oracle/build_ref.py)')
    import pandas as pd
    pd.set_option('future.infer_string', False)
    for m in ['netCDF4', 'osgeo', 'osgeo.ogr', 'osgeo.gdal', 'pathos',
              'pathos.multiprocessing', 'shapefile', 'cftime', 'matplotlib',
              'matplotlib.pyplot', 'descartes', 'pyximport']:
        sys.modules.setdefault(m, MagicMock())
    if 'spinterps' in sys.modules and not str(getattr(
            sys.modules['spinterps'], '__path__', [''])[0]).startswith(str(PKG)):
        raise RuntimeError('another package named spinterps is already imported')
    for name, sub in (('spinterps', ''), ('spinterps.cyth', 'cyth'), ('spinterps.interp', 'interp')):
        pkg = types.ModuleType(name)
        pkg.__path__ = [str(PKG / sub) if sub else str(PKG)]
        pkg.__package__ = name
        sys.modules[name] = pkg
    im = importlib.import_module('spinterps.cyth.interpmthds')
    cy = sys.modules['spinterps.cyth']
    for k in dir(im):
        if not k.startswith('_'):
            setattr(cy, k, getattr(im, k))           # what cyth/__init__.py re-exports
    misc = importlib.import_module('spinterps.misc')
    steps = importlib.import_module('spinterps.interp.steps')
    _loaded = (im, steps.SpInterpSteps, misc)
    return _loaded


class FakeLock:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run_case(case, intrp_dtype=np.float64, SpInterpSteps=None):
    """case: dict of plain arrays / settings (tests/golden/make_golden.py:base_case).
    Returns {label: ndarray[T, rows * cols]}, element 7 of the reference's 13-tuple
    (interp/steps.py:864-877)."""
    import pandas as pd
    if SpInterpSteps is None:
        SpInterpSteps = load()[1]
    n_stn = case['stn_xs'].size
    labels = [f'S{i:05d}' for i in range(n_stn)]
    T = case['data'].shape[0]
    tidx = pd.date_range('2000-01-01', periods=T)
    data_df = pd.DataFrame(case['data'].copy(), index=tidx, columns=labels)
    crds_df = pd.DataFrame({'X': case['stn_xs'], 'Y': case['stn_ys']}, index=labels)

    main = types.SimpleNamespace(
        _vb=False, _n_cpus=1, _mp_flag=False, _crds_df=crds_df,
        _min_var_thr=case.get('min_var_thr', -np.inf), _min_var_cut=case.get('min_var_cut'),
        _max_var_cut=case.get('max_var_cut'), _cntn_idxs=case.get('cntn_idxs'),
        _interp_crds_orig_shape=tuple(case['grid_shape']),
        _interp_x_crds_msh=case['cell_xs'].copy(),
        _interp_y_crds_msh=case['cell_ys'].copy(),
        _nc_file_path=None, _nc_nmrl_prcn=2,
        _neb_sel_mthd=case.get('neb_sel_mthd', 'all'), _n_nebs=case.get('n_nebs'),
        _n_pies=case.get('n_pies'),
        _min_vg_val=case.get('min_vg_val', 0.0),
        _interp_flag_est_vars=case.get('est_var_flag', False), _intrp_dtype=intrp_dtype)

    vgs_ser = None
    rord = None
    if case.get('vgs') is not None:
        vgs_ser = pd.Series(list(case['vgs']), index=tidx, dtype=object)
        rord = pd.Series(np.arange(T), index=tidx)
    stns_drft_df = None
    if case.get('stns_drft') is not None:
        stns_drft_df = pd.DataFrame(case['stns_drft'], index=labels)

    fld_end_row = case.get('fld_end_row')
    if fld_end_row is None:
        fld_end_row = case['grid_shape'][0]
    args = (data_df, 0, T, 1, case['interp_args'], FakeLock(), case.get('drft_arrs'),
            stns_drft_df, vgs_ser, rord, case.get('fld_beg_row', 0), fld_end_row)
    out = SpInterpSteps(main)._get_all_interp_outputs(args)     # fresh instance (quirk Q10)
    assert out is not None, 'the reference swallowed an exception (traceback above)'
    return out[7]
