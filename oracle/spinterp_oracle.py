"""CPU oracle for the spinterps gridded-interpolation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``spinterps_b200/`` may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or the timed CPU baseline, never as part of the product path.

It is a NumPy restatement (no pandas, no Cython) of what the reference does
between ``SpInterpSteps._get_all_interp_outputs`` (interp/steps.py:478-877) and
the Cython free functions of cyth/interpmthds.pyx.  Every function cites the
reference lines it follows.  Parity status: PINNED -- the restatement is
checked in ``tests/test_oracle_golden.py`` against arrays produced by running
the unmodified reference in the build container (``tests/golden/make_golden.py``,
fixtures ``tests/golden/*.npz``) and against the known-answer vectors listed in
SURVEY.md section 8c.

Two execution modes give the same numbers:
  * ``faithful=True``  keeps the reference's loop nest (per cell, per step;
    ``np.matmul`` + ``np.isclose`` per cell, steps.py:415-434).  This is the one
    timed as the CPU baseline because it has the reference's cost structure.
  * ``faithful=False`` vectorises the per-cell loops (same operations, batched)
    so that parity tests on bigger problems finish in seconds.
"""
from __future__ import annotations

import math

import numpy as np

# ---------------------------------------------------------------------------
# Variogram functions -- cyth/interpmthds.pyx:38-95
# ---------------------------------------------------------------------------


def _rng_vg(h, r, s):  # pyx:38-39
    return np.array(h, dtype=np.float64, copy=True)


def _nug_vg(h, r, s):  # pyx:42-43  (sill for EVERY h, including 0: quirk Q1)
    return np.full_like(h, s, dtype=np.float64)


def _sph_vg(h, r, s):  # pyx:46-55  (h >= r -> sill)
    a = (1.5 * h) / r
    b = (h * h * h) / (2 * (r * r * r))
    return np.where(h >= r, s, s * (a - b))


def _exp_vg(h, r, s):  # pyx:58-59
    return s * (1 - np.exp(-3 * h / r))


def _lin_vg(h, r, s):  # pyx:62-66  (h > r -> sill)
    return np.where(h > r, s, s * (h / r))


def _gau_vg(h, r, s):  # pyx:69-70
    return s * (1 - np.exp(-3 * ((h * h) / (r * r))))


def _pow_vg(h, r, s):  # pyx:73-74
    return s * np.power(h, r)


def _hol_vg(h, r, s):  # pyx:77-83
    a = (math.pi * h) / r
    with np.errstate(invalid='ignore', divide='ignore'):
        v = s * (1 - (np.sin(a) / a))
    return np.where(h == 0, 0.0, v)


ALL_VG_FTNS = {
    'Rng': _rng_vg, 'Nug': _nug_vg, 'Sph': _sph_vg, 'Exp': _exp_vg,
    'Lin': _lin_vg, 'Gau': _gau_vg, 'Pow': _pow_vg, 'Hol': _hol_vg}


def parse_vg_str(vg_models_str, clamp_range=True):
    """Split a variogram string into (sill, name, range) terms.

    Grammar as parsed at pyx:174-184: split on '+', strip, split on one space,
    then on '(' and ')'.  ``range = max(1e-5, float(range))`` (pyx:183) when
    ``clamp_range``.
    """
    terms = []
    for vg_model in str(vg_models_str).split('+'):
        vg_model = vg_model.strip()
        sill_s, vg_s = vg_model.split(' ')
        vg_s, range_s = vg_s.split('(')
        range_s = range_s.split(')')[0]
        rng = float(range_s)
        if clamp_range:
            rng = max(1e-5, rng)
        terms.append((float(sill_s), vg_s, rng))
    return terms


def fill_theo_vg_vals(vg_str, h_arr, r, s, vg_arr):
    """pyx:98-120 -- accumulates one variogram term into ``vg_arr``."""
    assert h_arr.shape[0]
    assert h_arr.shape[0] == vg_arr.shape[0]
    assert s >= 0
    assert r >= 0
    with np.errstate(invalid='ignore', divide='ignore'):
        vg_arr += ALL_VG_FTNS[vg_str.strip()](np.asarray(h_arr, dtype=np.float64), r, s)
    return


def get_theo_vg_vals(in_model, h_arr):
    """misc.py:1027-1047 (no range clamp on this route)."""
    vg_vals = np.zeros_like(h_arr, dtype=np.float64)
    for sill, name, rng in parse_vg_str(in_model, clamp_range=False):
        fill_theo_vg_vals(name, h_arr, rng, sill, vg_vals)
    return vg_vals


# ---------------------------------------------------------------------------
# Cython free functions -- cyth/interpmthds.pyx
# ---------------------------------------------------------------------------


def fill_dists_2d_mat(x1s, y1s, x2s, y2s, dists):
    """pyx:123-143: dists[i, j] = ((x1_i-x2_j)**2 + (y1_i-y2_j)**2)**0.5."""
    dx = x1s[:, None] - x2s[None, :]
    dy = y1s[:, None] - y2s[None, :]
    dists[...] = ((dx ** 2) + (dy ** 2)) ** 0.5
    return


def fill_dists_one_pt(x, y, xs, ys, dists):
    """pyx:768-781."""
    dists[...] = (((x - xs) ** 2) + ((y - ys) ** 2)) ** 0.5
    return


def fill_vg_var_arr(dists, in_vars, covar_flag, diag_mat_flag, vg_models_str, min_vg_val):
    """pyx:146-226.

    covar_flag=0 -> gamma(h); covar_flag=1 -> sum(sill) - gamma(h) accumulated
    term by term; values <= min_vg_val set to 0 afterwards; with diag_mat_flag
    the upper triangle INCLUDING the diagonal is computed and mirrored.
    """
    if covar_flag:
        cov_sign, cov_mult = -1, +1
    else:
        cov_sign, cov_mult = +1, +0

    in_vars[...] = 0.0
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        for sill_f, vg_s, range_f in parse_vg_str(vg_models_str):
            in_vars += (cov_mult * sill_f) + (cov_sign * ALL_VG_FTNS[vg_s](dists, range_f, sill_f))

        in_vars[in_vars <= min_vg_val] = 0.0

    if diag_mat_flag:
        # Upper triangle (g >= h) is authoritative; lower is its mirror.
        iu = np.triu_indices(in_vars.shape[0], 0, in_vars.shape[1])
        upper = np.zeros_like(in_vars)
        upper[iu] = in_vars[iu]
        mirrored = upper + np.triu(upper, 1).T
        in_vars[...] = mirrored
    return


def copy_2d_arr_at_idxs(arr, row_idxs, col_idxs, subset_arr):
    """pyx:229-248."""
    subset_arr[:row_idxs.shape[0], :col_idxs.shape[0]] = arr[np.ix_(row_idxs, col_idxs)]
    return


def fill_wts_and_sum(dists, wts, idw_exp):
    """pyx:784-795: w_i = 1/d_i**p, sequential sum."""
    wts_sum = 0.0
    with np.errstate(divide='ignore'):
        wts[...] = 1.0 / (dists ** idw_exp)
    for w in wts:  # sequential order as the C loop
        wts_sum += float(w)
    return wts_sum


def get_mults_sum(wts, data):
    """pyx:798-808: sequential dot product."""
    mults_sum = 0.0
    for w, z in zip(wts, data):
        mults_sum += float(w) * float(z)
    return mults_sum


# ---------------------------------------------------------------------------
# Host-side index logic
# ---------------------------------------------------------------------------


def check_full_nuggetness(in_model, min_vg_val):
    """misc.py:1074-1105."""
    in_model = str(in_model)
    nuggetness = False
    if in_model == 'nan':
        return nuggetness
    models = in_model.split('+')
    Sill = 0.0
    Range = 0.0
    for submodel in models:
        submodel = submodel.strip()
        Sill += float(submodel.split('(')[0].strip()[:-3].strip())
        Range = max(Range, float(submodel.split('(')[1].split(')')[0]))
    if (Sill <= min_vg_val) or (Range <= min_vg_val):
        nuggetness = True
    # misc.py:1102 compares the un-split string with 'Nug' and can never fire.
    return nuggetness


def get_vgs_cluster(vgs):
    """interp/vgclus.py:33-79: unique strings in first-occurrence order ->
    positions of the steps carrying them."""
    clus = {}
    for i, vg in enumerate(vgs):
        clus.setdefault(vg, []).append(i)
    return {k: np.asarray(v, dtype=np.int64) for k, v in clus.items()}


def get_grps_in_time(data):
    """interp/grps.py:57-101 on a [T, N] array (NaN = missing).

    Returns [(station index array, bool mask over steps)] in first-occurrence
    order of each distinct availability pattern.
    """
    avail = ~np.isnan(data)
    keys = [row.tobytes() for row in avail]
    seen = {}
    grps = []
    for i, key in enumerate(keys):
        if key in seen:
            grps[seen[key]][1][i] = True
            continue
        seen[key] = len(grps)
        mask = np.zeros(data.shape[0], dtype=bool)
        mask[i] = True
        grps.append((np.where(avail[i])[0], mask))
    return grps


def _get_neb_idxs_grps(all_neb_idxs):
    """interp/grps.py:103-139: cells with identical neighbour rows, groups in
    first-occurrence order, members ascending."""
    seen = {}
    grps = []
    for i in range(all_neb_idxs.shape[0]):
        key = all_neb_idxs[i].tobytes()
        if key in seen:
            grps[seen[key]].append(i)
        else:
            seen[key] = len(grps)
            grps.append([i])
    return [np.asarray(g, dtype=np.int64) for g in grps]


def get_neb_idxs_and_grps(neb_sel_mthd, n_nebs, dst_xs, dst_ys, ref_xs, ref_ys, n_pies=None):
    """interp/grps.py:249-288 ('all' :141-145, 'nrst' :147-166, 'pie' :168-247)."""
    n_refs = ref_xs.size
    n_dst = dst_xs.shape[0]
    if neb_sel_mthd == 'all':
        all_neb_idxs = np.tile(np.arange(n_refs), (n_dst, 1))
        # every row identical -> one group holding every cell
        return all_neb_idxs, [np.arange(n_dst, dtype=np.int64)]

    if neb_sel_mthd == 'nrst':
        all_neb_idxs = np.full((n_dst, min(n_nebs, n_refs)), -1, dtype=int)
        for i in range(n_dst):
            dists = (((dst_xs[i] - ref_xs) ** 2) + ((dst_ys[i] - ref_ys) ** 2)) ** 0.5
            all_neb_idxs[i, :] = np.sort(np.argsort(dists)[:n_nebs])
        return all_neb_idxs, _get_neb_idxs_grps(all_neb_idxs)

    if neb_sel_mthd == 'pie':
        return (lambda a: (a, _get_neb_idxs_grps(a)))(
            _get_pie_neb_idxs(n_nebs, n_pies, dst_xs, dst_ys, ref_xs, ref_ys))

    raise NotImplementedError(neb_sel_mthd)


def sel_equidist_refs(dst_x, dst_y, ref_xs, ref_ys, n_pies, min_dist_thresh, not_neb_flag,
                      dists, tem_ref_sel_dists, ref_sel_pie_idxs, ref_pie_idxs, ref_pie_cts):
    """cyth/interpmthds.pyx:811-890.  Angular sector of every reference point seen
    from the destination, and its distance rank INSIDE its sector
    (``ref_sel_pie_idxs``).  A point within ``min_dist_thresh`` short-cuts to rank 0
    for the nearest point only (the caller always passes -1: never)."""
    n_refs = ref_xs.size
    two_pi = 2 * math.pi
    ref_sel_pie_idxs[:] = not_neb_flag
    fill_dists_one_pt(dst_x, dst_y, ref_xs, ref_ys, dists)
    min_dist, min_dist_idx = np.inf, -1
    for i in range(n_refs):
        if (dists[i] <= min_dist_thresh) and (dists[i] < min_dist):
            min_dist_idx, min_dist = i, dists[i]
    if min_dist < np.inf:
        ref_sel_pie_idxs[min_dist_idx] = 0
        return
    ref_pie_cts[:] = 0
    for j in range(n_refs):
        x_dist = ref_xs[j] - dst_x
        y_dist = ref_ys[j] - dst_y
        if not x_dist:
            ang = 0.0
        else:
            ang = math.atan(y_dist / x_dist)
            if (x_dist < 0) and (y_dist > 0):
                ang = math.pi + ang
            elif (x_dist < 0) and (y_dist < 0):
                ang = math.pi + ang
            elif (x_dist > 0) and (y_dist < 0):
                ang = two_pi + ang
        pie = int(ang * n_pies / two_pi)   # x_dist < 0, y_dist == 0 gives -0.0 -> sector 0
        ref_pie_idxs[j] = pie
        ref_pie_cts[pie] += 1
    for j in range(n_pies):
        if not ref_pie_cts[j]:
            continue
        tem_ref_sel_dists[:] = np.where(np.asarray(ref_pie_idxs) == j, dists, np.inf)
        srtd = np.argsort(tem_ref_sel_dists)
        for i in range(n_refs):
            if tem_ref_sel_dists[srtd[i]] == np.inf:
                break
            ref_sel_pie_idxs[srtd[i]] = i


def get_nd_dists(pts):
    """cyth/interpmthds.pyx:893-925: distances of all point pairs i > j, row by row."""
    pts = np.asarray(pts, dtype=np.float64)
    n_pts = pts.shape[0]
    out = np.full((n_pts * (n_pts - 1)) // 2, np.nan)
    c = 0
    for i in range(n_pts):
        for j in range(i):
            d = 0.0
            for k in range(pts.shape[1]):
                d += (pts[i, k] - pts[j, k]) ** 2
            out[c] = d ** 0.5
            c += 1
    return out


def _get_pie_neb_idxs(n_nebs, n_pies, dst_xs, dst_ys, ref_xs, ref_ys):
    """interp/grps.py:168-247: stations ordered by (rank inside their sector, distance);
    the first n_nebs of that order, ascending.  (The reference passes uint32 work arrays
    where the compiled helper wants unsigned long, so this only runs on LLP64; the golden
    case h_pie was produced with a cast at that call boundary.)"""
    n_refs = ref_xs.size
    dists = np.zeros(n_refs)
    ref_pie_idxs = np.zeros(n_refs, dtype=np.int64)
    tem = np.zeros(n_refs)
    sel = np.zeros(n_refs, dtype=np.int64)
    cts = np.zeros(n_pies, dtype=np.int64)
    all_neb_idxs = np.full((dst_xs.size, n_nebs), -1, dtype=int)
    for i in range(dst_xs.size):
        sel_equidist_refs(dst_xs[i], dst_ys[i], ref_xs, ref_ys, n_pies, -1, -1, dists, tem,
                          sel, ref_pie_idxs, cts)
        assert np.all(sel != -1)
        order = []
        for u in np.unique(sel):
            same = np.where(sel == u)[0]
            order.extend(same[np.argsort(dists[same])].tolist())
        all_neb_idxs[i, :] = np.sort(np.array(order)[:n_nebs])
    return all_neb_idxs


# ---------------------------------------------------------------------------
# Kriging system assembly -- interp/steps.py:170-243
# ---------------------------------------------------------------------------


def get_vars_arr_subset(interp_type, ref_drfts, ref_ref_vars_all, dst_ref_vars_all,
                        ref_ref_sub_idxs, cell_idxs, dst_drfts):
    if interp_type == 'OK':
        add_rc = 1
    elif interp_type == 'SK':
        add_rc = 0
    elif interp_type == 'EDK':
        add_rc = 1 + ref_drfts.shape[1]
    else:
        raise NotImplementedError

    n = ref_ref_sub_idxs.size
    A = np.full((n + add_rc, n + add_rc), np.nan)
    copy_2d_arr_at_idxs(ref_ref_vars_all, ref_ref_sub_idxs, ref_ref_sub_idxs, A)
    R = np.full((cell_idxs.size, n + add_rc), np.nan)
    copy_2d_arr_at_idxs(dst_ref_vars_all, cell_idxs, ref_ref_sub_idxs, R)

    if interp_type == 'OK':  # steps.py:212-216
        A[n, :n] = 1.0
        A[:, n] = 1.0
        A[n, n] = 0.0
        R[:, n] = 1.0
    elif interp_type == 'EDK':  # steps.py:221-234
        A[n, :n] = 1.0
        A[:n, n] = 1.0
        R[:, n] = 1.0
        A[n:, n:] = 0.0
        for k in range(ref_drfts.shape[1]):
            A[n + 1 + k, :n] = ref_drfts[:, k]
            A[:n, n + 1 + k] = ref_drfts[:, k]
            R[:, n + 1 + k] = dst_drfts[k, :]
    return A, R


# ---------------------------------------------------------------------------
# _get_interp -- interp/steps.py:245-401 (+ :403-435)
# ---------------------------------------------------------------------------


def _krige_fill_faithful(n_refs, inv, R, dists_sub, ref_data_j, dst_row, est_row):
    """steps.py:415-434, one step, python loop over cells."""
    for i in range(R.shape[0]):
        lmds = np.matmul(inv, R[i])
        if not np.isclose(lmds[:n_refs].sum(), 1.0):
            dst_row[i] = ref_data_j[np.argmin(dists_sub[i])]
            if est_row is not None:
                est_row[i] = 0.0
        else:
            dst_row[i] = (lmds[:n_refs] * ref_data_j).sum()
            if est_row is not None:
                est_row[i] = (lmds * R[i]).sum() + lmds[n_refs]
    return


def _krige_fill_fast(n_refs, lmds_all, lsum_ok, nnb_idx, R, ref_data_j, dst_row, est_row):
    """Vectorised equivalent of steps.py:415-434 (weights precomputed once per
    system because they do not depend on the step)."""
    vals = lmds_all[:, :n_refs] @ ref_data_j
    nn = ref_data_j[nnb_idx]
    dst_row[...] = np.where(lsum_ok, vals, nn)
    if est_row is not None:
        ev = (lmds_all * R).sum(axis=1) + lmds_all[:, n_refs]
        est_row[...] = np.where(lsum_ok, ev, 0.0)
    return


def get_interp(ref_data, interp_type, ref_drfts, dst_drfts, idw_exp, models,
               interp_steps_flags, nuggetness_flags, dists_sub,
               ref_ref_vars_all, dst_ref_vars_all, ref_ref_sub_idxs, cell_idxs,
               est_var_flag, est_dtype, prblm_steps, faithful):
    """interp/steps.py:245-401.  ``dists_sub`` is modified in place by IDW
    (steps.py:297-301, quirk Q7) exactly like the reference."""
    n_dsts = dists_sub.shape[0]
    n_time, n_refs = ref_data.shape

    dst_data = np.full((n_time, n_dsts), np.nan)
    est_vars = None
    if est_var_flag and interp_type == 'OK':
        est_vars = np.full((n_time, n_dsts), np.nan, dtype=est_dtype)

    ref_means = ref_data.mean(axis=1)

    if n_refs == 1:  # steps.py:282-283
        for j in range(n_time):
            dst_data[j, :] = ref_data[j, :]

    elif interp_type == 'NNB':  # steps.py:285-291
        nnb_idx = np.argmin(dists_sub, axis=1)
        dst_data[...] = ref_data[:, nnb_idx]

    elif interp_type == 'IDW':  # steps.py:293-313
        if faithful:
            wts = np.full(n_refs, np.nan)
            for i in range(n_dsts):
                dists = dists_sub[i]
                dists_max = dists.max()
                if dists_max > 0:
                    dists /= dists_max
                wts_sum = fill_wts_and_sum(dists, wts, idw_exp)
                assert wts_sum >= 1e-14, wts_sum
                for j in range(n_time):
                    if interp_steps_flags[j]:
                        dst_data[j, i] = get_mults_sum(wts, ref_data[j]) / wts_sum
                    else:
                        dst_data[j, i] = ref_means[j]
        else:
            dmax = dists_sub.max(axis=1, keepdims=True)
            np.divide(dists_sub, dmax, out=dists_sub, where=dmax > 0)
            with np.errstate(divide='ignore', invalid='ignore'):
                wts = 1.0 / (dists_sub ** idw_exp)
                wts_sum = np.cumsum(wts, axis=1)[:, -1]  # sequential order
                assert np.all(~(wts_sum < 1e-14)), wts_sum.min()
                for j in range(n_time):
                    if interp_steps_flags[j]:
                        ms = np.cumsum(wts * ref_data[j][None, :], axis=1)[:, -1]
                        dst_data[j, :] = ms / wts_sum
                    else:
                        dst_data[j, :] = ref_means[j]

    elif interp_type in ('OK', 'SK', 'EDK'):  # steps.py:315-396
        old_model = ''
        last_failed_flag = False
        covar_flag = 1 if interp_type == 'SK' else 0
        nnb_idx = None
        for j in np.argsort(models):
            model = models[j]
            if (not interp_steps_flags[j]) or nuggetness_flags[j]:
                dst_data[j, :] = ref_means[j]
                if est_vars is not None:
                    est_vars[j, :] = 0.0
                continue
            if model == 'nan':
                raise ValueError('NaN VG!')
            if model != old_model:
                A, R = get_vars_arr_subset(
                    interp_type, ref_drfts,
                    ref_ref_vars_all[(model, covar_flag)],
                    dst_ref_vars_all[(model, covar_flag)],
                    ref_ref_sub_idxs, cell_idxs, dst_drfts)
                old_model = model
                try:
                    inv = np.linalg.pinv(A)
                    last_failed_flag = False
                except Exception:
                    last_failed_flag = True
                    if j not in prblm_steps:
                        prblm_steps.append(j)
                    inv = np.nan
                if (not faithful) and (not last_failed_flag):
                    lmds_all = R @ inv.T
                    with np.errstate(invalid='ignore'):
                        lsum_ok = np.isclose(lmds_all[:, :n_refs].sum(axis=1), 1.0)

            if (not faithful) and nnb_idx is None:
                nnb_idx = np.argmin(dists_sub, axis=1)

            if last_failed_flag:  # steps.py:372-384
                if faithful:
                    for i in range(n_dsts):
                        dst_data[j, i] = ref_data[j, np.argmin(dists_sub[i])]
                else:
                    dst_data[j, :] = ref_data[j, nnb_idx]
                if est_vars is not None:
                    est_vars[j, :] = 0.0
            elif faithful:
                _krige_fill_faithful(
                    n_refs, inv, R, dists_sub, ref_data[j], dst_data[j],
                    None if est_vars is None else est_vars[j])
            else:
                _krige_fill_fast(
                    n_refs, lmds_all, lsum_ok, nnb_idx, R, ref_data[j], dst_data[j],
                    None if est_vars is None else est_vars[j])
    else:
        raise NotImplementedError(interp_type)

    return dst_data, est_vars


# ---------------------------------------------------------------------------
# _get_all_interp_outputs -- interp/steps.py:478-877
# ---------------------------------------------------------------------------


def interp_chunk(
        data, stn_xs, stn_ys, cell_xs, cell_ys, grid_shape, interp_args,
        vgs=None, cntn_idxs=None, drft_arrs=None, stns_drft=None,
        fld_beg_row=0, fld_end_row=None,
        neb_sel_mthd='all', n_nebs=None,
        min_var_thr=-np.inf, min_var_cut=None, max_var_cut=None,
        min_vg_val=0.0, est_var_flag=False, intrp_dtype=np.float32,
        faithful=False, n_pies=None):
    """Compute half of ``SpInterpSteps.interpolate_subset`` for one time chunk
    and one grid-row chunk.

    data [T, N] f64 (NaN = missing); stn_xs/ys [N]; cell_xs/ys = the (masked)
    raveled cell-centre coordinates of the WHOLE grid (``_interp_x/y_crds_msh``);
    grid_shape = ``_interp_crds_orig_shape``; interp_args = list of
    (type, fig_dir, label[, idw_exp]) in the reference's order; vgs = list of T
    variogram strings or None; cntn_idxs = bool mask over the raveled grid or
    None; drft_arrs [n_drifts, cells]; stns_drft [N, n_drifts].

    Returns {label: ndarray[T, rows*cols] of intrp_dtype} (element 7 of the
    reference's 13-tuple, steps.py:864-877) and the list of problem steps.
    """
    data = np.asarray(data, dtype=np.float64)
    stn_xs = np.asarray(stn_xs, dtype=np.float64)
    stn_ys = np.asarray(stn_ys, dtype=np.float64)
    if fld_end_row is None:
        fld_end_row = grid_shape[0]

    interp_types = [a[0] for a in interp_args]
    interp_labels = [a[2] for a in interp_args]
    krg_flag = any(t in interp_types for t in ('OK', 'SK', 'EDK'))
    edk_flag = 'EDK' in interp_types
    if krg_flag:
        assert all(vg != 'nan' for vg in vgs), 'NaN VGs not allowed!'

    n_steps = data.shape[0]

    # -- cell subsetting, steps.py:512-568 --
    fld_n_cols = grid_shape[1]
    fld_beg_idx = fld_beg_row * fld_n_cols
    fld_end_idx = fld_end_row * fld_n_cols
    fld_grd_shape = ((fld_end_row - fld_beg_row), fld_n_cols)

    if cntn_idxs is not None:
        cntn_idxs_whr = np.where(cntn_idxs)[0]
        sel = (cntn_idxs_whr >= fld_beg_idx) & (cntn_idxs_whr < fld_end_idx)
        msh_idxs = np.arange(cntn_idxs_whr.size)[sel]
        cntn_idxs_whr = cntn_idxs_whr[sel] - fld_beg_idx
        dst_xs = cell_xs[msh_idxs]
        dst_ys = cell_ys[msh_idxs]
        if drft_arrs is not None:
            drft_arrs = drft_arrs[:, msh_idxs]
    else:
        cntn_idxs_whr = None
        dst_xs = cell_xs[fld_beg_idx:fld_end_idx]
        dst_ys = cell_ys[fld_beg_idx:fld_end_idx]
        if drft_arrs is not None:
            drft_arrs = drft_arrs[:, fld_beg_idx:fld_end_idx]
    n_dst_pts = dst_xs.shape[0]

    if vgs is not None:
        vgs = list(vgs)
        vgs_clus = get_vgs_cluster(vgs)

    # -- drop stations never selected, steps.py:592-608 (Q9) --
    tke = np.unique(get_neb_idxs_and_grps(
        neb_sel_mthd, n_nebs, dst_xs, dst_ys, stn_xs, stn_ys, n_pies)[0])
    if tke.size != stn_xs.shape[0]:
        stn_xs = stn_xs[tke].copy()
        stn_ys = stn_ys[tke].copy()
        data = data[:, tke].copy()
        if stns_drft is not None:
            stns_drft = stns_drft[tke, :]

    grps_in_time = get_grps_in_time(data)

    # -- distance and variogram matrices, steps.py:620-653 --
    def svars(dists, diag):
        out = {}
        var_flag = any(t in interp_types for t in ('OK', 'EDK'))
        for vg in vgs_clus.keys():
            if var_flag:
                arr = np.full_like(dists, np.nan)
                fill_vg_var_arr(dists, arr, 0, diag, vg, min_vg_val)
                out[(vg, 0)] = arr
            if 'SK' in interp_types:
                arr = np.full_like(dists, np.nan)
                fill_vg_var_arr(dists, arr, 1, diag, vg, min_vg_val)
                out[(vg, 1)] = arr
        return out

    if vgs is not None:
        rr = np.full((stn_xs.size, stn_xs.size), np.nan)
        fill_dists_2d_mat(stn_xs, stn_ys, stn_xs, stn_ys, rr)
        ref_ref_vars_all = svars(rr, 1)
    else:
        ref_ref_vars_all = None

    dst_ref_dists_all = np.full((n_dst_pts, stn_xs.size), np.nan)
    fill_dists_2d_mat(dst_xs, dst_ys, stn_xs, stn_ys, dst_ref_dists_all)
    dst_ref_vars_all = svars(dst_ref_dists_all, 0) if vgs is not None else None

    flds = {lab: np.full((n_steps, int(np.prod(fld_grd_shape))), np.nan, dtype=intrp_dtype)
            for lab in interp_labels}

    prblm_steps = []

    for stn_idxs, step_mask in grps_in_time:  # HOT LOOP A, steps.py:673
        step_idxs = np.where(step_mask)[0]
        if not stn_idxs.size:  # steps.py:677-688 (Q11)
            for s in step_idxs:
                if s not in prblm_steps:
                    prblm_steps.append(int(s))
            continue

        grp_xs = stn_xs[stn_idxs]
        grp_ys = stn_ys[stn_idxs]
        grp_data = data[np.ix_(step_idxs, stn_idxs)]
        grp_drifts = stns_drft[stn_idxs] if edk_flag else None

        if krg_flag:
            vg_models = np.array([vgs[s] for s in step_idxs], dtype=object)
            nuggetness_flags = np.array(
                [check_full_nuggetness(vg, min_vg_val) for vg in vg_models], dtype=bool)
        else:
            vg_models = None
            nuggetness_flags = None

        neb_idxs, neb_grps = get_neb_idxs_and_grps(
            neb_sel_mthd, n_nebs, dst_xs, dst_ys, grp_xs, grp_ys, n_pies)

        pts_done = np.zeros(n_dst_pts, dtype=bool)

        for cell_grp in neb_grps:  # HOT LOOP B, steps.py:740
            sub_ref_idxs = neb_idxs[cell_grp[0]]
            sub_ref_data = grp_data[:, sub_ref_idxs].copy('C')
            ref_ref_sub_idxs = stn_idxs[sub_ref_idxs]

            dists_sub = np.full((cell_grp.size, ref_ref_sub_idxs.size), np.nan)
            copy_2d_arr_at_idxs(dst_ref_dists_all, cell_grp, ref_ref_sub_idxs, dists_sub)

            # steps.py:760-765
            with np.errstate(invalid='ignore'):
                interp_steps_flags = np.any(sub_ref_data >= min_var_thr, axis=1)

            if edk_flag:
                sub_ref_drifts = grp_drifts[sub_ref_idxs, :]
                sub_dst_drifts = drft_arrs[:, cell_grp]
            else:
                sub_ref_drifts = sub_dst_drifts = None

            for i, interp_type in enumerate(interp_types):
                if interp_labels[i] == 'EST_VARS_OK':
                    continue
                idw_exp = interp_args[i][3] if interp_type == 'IDW' else 5

                vals, est_vars = get_interp(
                    sub_ref_data, interp_type, sub_ref_drifts, sub_dst_drifts,
                    idw_exp, vg_models, interp_steps_flags, nuggetness_flags,
                    dists_sub, ref_ref_vars_all, dst_ref_vars_all,
                    ref_ref_sub_idxs, cell_grp, est_var_flag, intrp_dtype,
                    prblm_steps, faithful)

                # _mod_min_max, steps.py:466-476
                with np.errstate(invalid='ignore'):
                    if min_var_cut is not None:
                        vals[vals < min_var_cut] = min_var_cut
                    if max_var_cut is not None:
                        vals[vals > max_var_cut] = max_var_cut

                out_pos = cell_grp if cntn_idxs_whr is None else cntn_idxs_whr[cell_grp]
                fld = flds[interp_labels[i]]
                fld[np.ix_(step_idxs, out_pos)] = vals.astype(intrp_dtype)
                if est_vars is not None:
                    flds['EST_VARS_OK'][np.ix_(step_idxs, out_pos)] = est_vars

            pts_done[cell_grp] = True
        assert np.all(pts_done), 'Some points not interpolated!'

    return flds, prblm_steps


# ---------------------------------------------------------------------------
# Grid preparation.  The reference does the geometric part with OGR /
# GDAL (misc.py:407-540 ``chk_pt_cntmnt_in_polys_mp`` = Contains on polygons buffered with
# geom.Buffer, interp/drift.py:25-226), which are absent from the build container.
# Parity is pinned as far as the reference's own code goes: ``tests/golden/p*_prep_*.npz``
# are outputs of the reference's preparation methods run unmodified over in-memory
# stand-ins for the two libraries (``tests/golden/make_golden_prep.py``) -- bounds, row /
# column window, cell centres, the containment search around the predicate, the cell mask,
# the selected stations, drift values of cells and stations.  NOT pinned: GEOS's own
# predicates on points closer to an outline than the sagitta of its buffer arcs (30
# segments per quadrant) -- the fixtures hold no such point, and here the exact distance is
# used.  The functions below are the checker for the CUDA kernels: the same formulas in
# plain NumPy (IEEE operations in the same order, no FMA).
# ---------------------------------------------------------------------------
def aligned_grid_bounds(align, extent, cell_bdist):
    """misc.py:743-885 ``get_aligned_shp_bds_and_cell_size``: the extent (+- cell_bdist) of
    the polygons -- or, without polygons, of the alignment raster itself -- moved outwards
    to the raster's cell lattice.  align = (x_min, y_max, cell_size, n_rows, n_cols)."""
    rx_min, ry_max, cs, n_rows, n_cols = align
    rel_cell_err = 1e-5
    abs_cell_err = abs(cs * rel_cell_err)
    rx_max = rx_min + (n_cols * cs)                       # misc.py:657-658
    ry_min = ry_max - (n_rows * cs)
    if extent is not None:
        sx_min, sx_max, sy_min, sy_max = extent
        if cell_bdist:
            sx_min -= cell_bdist
            sx_max += cell_bdist
            sy_min -= cell_bdist
            sy_max += cell_bdist
    else:
        sx_min, sx_max, sy_min, sy_max = rx_min, rx_max, ry_min, ry_max
    if sx_min < rx_min:
        assert abs(sx_min - rx_min) <= abs_cell_err
    if sx_max > rx_max:
        assert abs(sx_max - rx_max) <= abs_cell_err
    if sy_min < ry_min:
        assert abs(sy_min - ry_min) <= abs_cell_err
    if sy_max > ry_max:
        assert abs(sy_max - ry_max) <= abs_cell_err
    if not np.isclose(sx_min, rx_min, rtol=0, atol=rel_cell_err):
        rem = ((sx_min - rx_min) / cs) % 1
        adj = rem * cs
        ax_min = sx_min - adj
    else:
        ax_min = rx_min
    if not np.isclose(sy_max, ry_max, rtol=0, atol=rel_cell_err):
        rem = ((ry_max - sy_max) / cs) % 1
        adj = rem * cs
        ay_max = sy_max + adj
    else:
        ay_max = ry_max
    if not np.isclose(sx_max, rx_max, rtol=0, atol=rel_cell_err):
        rem = ((sx_max - rx_min) / cs) % 1
        adj = rem * cs
        ax_max = sx_max + (cs - adj)
    else:
        ax_max = rx_max
    if not np.isclose(sy_min, ry_min, rtol=0, atol=rel_cell_err):
        rem = ((ry_max - sy_min) / cs) % 1
        adj = rem * cs
        ay_min = sy_min - (cs - adj)
    else:
        ay_min = ry_min
    assert ax_min >= rx_min and ax_max <= rx_max and ay_min >= ry_min and ay_max <= ry_max
    return ax_min, ax_max, ay_min, ay_max


def prepare_grid(stn_xs, stn_ys, cell_size, cell_bdist=0.0, rings=None, raster_geo=None,
                 align=None):
    """Bounds (interp/prepare.py:107-137, or :45-90 with an alignment raster), row / column
    window (:150-182) and cell-centre coordinates (:188-199) of the interpolation grid.

    rings : outer rings of the selection polygons or None (bounds from the stations).
    raster_geo : (x_min, y_max, n_rows, n_cols) of the drift rasters or None; their bounds
    are rounded to 6 decimals (interp/drift.py:134-152) and the window is relative to them.
    align : (x_min, y_max, cell_size, n_rows, n_cols) of the alignment raster or None.
    Returns (bounds [x_min, x_max, y_min, y_max], window [min_row, max_row, min_col,
    max_col], x coordinates [n_cols], y coordinates [n_rows])."""
    import math
    if rings is not None:
        allv = np.concatenate([np.asarray(r, dtype=np.float64) for r in rings], axis=0)
        x_min, x_max = allv[:, 0].min(), allv[:, 0].max()
        y_min, y_max = allv[:, 1].min(), allv[:, 1].max()
    else:
        x_min, x_max = np.min(stn_xs), np.max(stn_xs)
        y_min, y_max = np.min(stn_ys), np.max(stn_ys)
    if align is not None:
        x_min, x_max, y_min, y_max = aligned_grid_bounds(
            align, (float(x_min), float(x_max), float(y_min), float(y_max)) if rings is not None
            else None, cell_bdist)
        cell_bdist = 0.0
    x_min -= cell_bdist
    x_max += cell_bdist
    y_min -= cell_bdist
    y_max += cell_bdist
    if raster_geo is not None:
        rx_min, ry_max, n_rows, n_cols = raster_geo
        rx_min, rx_max, ry_min, ry_max = np.round(
            (rx_min, rx_min + n_cols * cell_size, ry_max - n_rows * cell_size, ry_max), 6)
        assert x_min >= rx_min and x_max <= rx_max and y_min >= ry_min and y_max <= ry_max
        min_col = int(math.floor((x_min - rx_min) / cell_size))
        max_col = int(math.ceil((x_max - rx_min) / cell_size)) - 1
        min_row = int(math.floor((ry_max - y_max) / cell_size))
        max_row = int(math.ceil((ry_max - y_min) / cell_size)) - 1
    else:
        min_col = min_row = 0
        max_col = int(math.ceil((x_max - x_min) / cell_size)) - 1
        max_row = int(math.ceil((y_max - y_min) / cell_size)) - 1
    assert 0 <= min_col <= max_col and 0 <= min_row <= max_row
    strt_x = x_min + (0.5 * cell_size)
    end_x = strt_x + ((max_col - min_col) * cell_size)
    strt_y = y_max - (0.5 * cell_size)
    end_y = strt_y - ((max_row - min_row) * cell_size)
    xs = np.linspace(strt_x, end_x, (max_col - min_col + 1))
    ys = np.linspace(strt_y, end_y, (max_row - min_row + 1))
    return (np.array([x_min, x_max, y_min, y_max]),
            np.array([min_row, max_row, min_col, max_col], dtype=np.int64), xs, ys)


def drift_window_indices(window, cntn_idxs=None):
    """Raster (row, col) of every (selected) grid cell, interp/drift.py:175-188."""
    min_row, max_row, min_col, max_col = (int(v) for v in window)
    cols, rows = np.meshgrid(np.arange(min_col, max_col + 1), np.arange(min_row, max_row + 1))
    rows, cols = rows.ravel(), cols.ravel()
    if cntn_idxs is not None:
        rows, cols = rows[cntn_idxs], cols[cntn_idxs]
    return rows, cols


def drift_station_indices(stn_xs, stn_ys, ras_x_min, ras_y_max, cell_size):
    """Raster (row, col) of the stations, interp/drift.py:209-210 (``int()`` truncates)."""
    cols = np.array([int((x - ras_x_min) / cell_size) for x in np.asarray(stn_xs)], dtype=np.int64)
    rows = np.array([int((ras_y_max - y) / cell_size) for y in np.asarray(stn_ys)], dtype=np.int64)
    return rows, cols


def points_in_polygons(xs, ys, rings, buffer_dist=0.0):
    """bool [n]: inside any ring by the even-odd crossing rule, or (buffer_dist > 0)
    closer than buffer_dist to a ring edge."""
    xs = np.asarray(xs, dtype=np.float64).ravel()
    ys = np.asarray(ys, dtype=np.float64).ravel()
    inside = np.zeros(xs.size, dtype=bool)
    buf2 = np.float64(buffer_dist) * np.float64(buffer_dist)
    for ring in rings:
        r = np.asarray(ring, dtype=np.float64)
        if r.shape[0] >= 2 and np.array_equal(r[0], r[-1]):
            r = r[:-1]
        nxt = np.roll(r, -1, axis=0)
        parity = np.zeros(xs.size, dtype=bool)
        for (ax, ay), (bx, by) in zip(r, nxt):
            dx, dy = bx - ax, by - ay
            strad = (ay > ys) != (by > ys)
            with np.errstate(divide='ignore', invalid='ignore'):
                xi = dx * (ys - ay) / dy + ax
            parity ^= strad & (xs < xi)
            if buffer_dist > 0:
                wx, wy = xs - ax, ys - ay
                l2 = dx * dx + dy * dy
                t = (wx * dx + wy * dy) / l2 if l2 > 0 else np.zeros_like(wx)
                t = np.minimum(np.maximum(t, 0.0), 1.0)
                qx, qy = wx - t * dx, wy - t * dy
                inside |= (qx * qx + qy * qy) < buf2
        inside |= parity
    return inside


def sample_raster(ras, rows, cols, ndv=None):
    """interp/drift.py:190-198, :209-221: raster values at (row, col), no-data -> NaN."""
    ras = np.asarray(ras, dtype=np.float64)
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    ok = (rows >= 0) & (rows < ras.shape[0]) & (cols >= 0) & (cols < ras.shape[1])
    out = np.full(rows.shape, np.nan)
    out[ok] = ras[rows[ok], cols[ok]]
    if ndv is not None:
        with np.errstate(invalid='ignore'):
            out[np.isclose(ndv, out)] = np.nan
    return out
