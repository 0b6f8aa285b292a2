"""Load the golden fixtures written by tests/golden/make_golden.py."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / 'golden'

CASES = ['a_ok_idw_nnb', 'b_ok_groups_flags', 'c_edk_drift', 'd_sk_ok_mask_rows',
         'e_nrst', 'f_vg_families', 'g_idw_only', 'h_pie']


def _opt(z, key, conv=None):
    if key not in z.files:
        return None
    v = z[key]
    if v.ndim == 0:
        v = v.item()
    return conv(v) if conv else v


def load_case(name):
    z = np.load(GOLDEN / f'{name}.npz', allow_pickle=False)
    interp_args = []
    for t, lab, e in zip(z['ia_types'], z['ia_labels'], z['ia_exps']):
        a = (str(t), None, str(lab))
        if not np.isnan(e):
            a = a + (float(e),)
        interp_args.append(a)
    case = dict(
        data=z['data'], stn_xs=z['stn_xs'], stn_ys=z['stn_ys'],
        cell_xs=z['cell_xs'], cell_ys=z['cell_ys'],
        grid_shape=tuple(int(v) for v in z['grid_shape']),
        interp_args=interp_args,
        vgs=[str(v) for v in z['vgs']] if 'vgs' in z.files else None,
        cntn_idxs=_opt(z, 'cntn_idxs'), drft_arrs=_opt(z, 'drft_arrs'),
        stns_drft=_opt(z, 'stns_drft'),
        fld_beg_row=int(z['fld_beg_row']), fld_end_row=int(z['fld_end_row']),
        neb_sel_mthd=str(z['neb_sel_mthd']), n_nebs=_opt(z, 'n_nebs', int),
        min_var_thr=float(z['min_var_thr']),
        min_var_cut=_opt(z, 'min_var_cut', float), max_var_cut=_opt(z, 'max_var_cut', float),
        min_vg_val=float(z['min_vg_val']), est_var_flag=bool(z['est_var_flag']))
    if 'n_pies' in z.files:
        case['n_pies'] = int(z['n_pies'])
    outs = {k[5:]: z[k] for k in z.files if k.startswith('out__')}
    return case, outs


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor) with NaN positions required to coincide."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), 'NaN pattern differs'
    if not (~nb).any():
        return 0.0
    d = np.abs(a[~nb] - b[~nb])
    return float((d / np.maximum(np.abs(b[~nb]), floor)).max())
