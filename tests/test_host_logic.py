"""Host-side logic (no GPU): grouping indices bit-exact vs the oracle, variogram
helpers, sharding, output-file layout, configuration surface."""
import numpy as np
import pandas as pd
import pytest

from oracle import spinterp_oracle as orc
from spinterps_b200 import dist as sdist
from spinterps_b200 import ncwriter, vgs
from spinterps_b200.engine import availability_groups
from spinterps_b200.main import SpInterpMain


@pytest.mark.parametrize('seed,T,N,miss', [(0, 50, 6, 0.4), (1, 300, 40, 0.1), (2, 20, 3, 0.9),
                                           (3, 7, 9, 0.0)])
def test_availability_groups_bit_exact(seed, T, N, miss):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(T, N))
    d[rng.random((T, N)) < miss] = np.nan
    grp_of_step, grp_mask = availability_groups(~np.isnan(d))
    ref = orc.get_grps_in_time(d)          # interp/grps.py:57-101 restated
    assert len(ref) == grp_mask.shape[0]
    for k, (idx, mask) in enumerate(ref):
        assert np.array_equal(np.where(grp_mask[k])[0], idx)
        assert np.array_equal(grp_of_step == k, mask)


def test_nuggetness_and_cluster_match_oracle():
    cases = ['0.0 Nug(0.0)', '0.1 Nug(0.0) + 0.9 Sph(20000)', '0.00001 Nug(0.0) + 0.00002 Sph(20000)',
             '1.0 Exp(0.00001)', '0.5 Sph(10)+0.5 Gau(20)', 'nan']
    for c in cases:
        for mv in (0.0, 1e-4, 0.5):
            assert vgs.check_full_nuggetness(c, mv) == orc.check_full_nuggetness(c, mv), (c, mv)
    seq = ['a', 'b', 'a', 'c', 'b', 'a']
    got, exp = vgs.get_vgs_cluster(seq), orc.get_vgs_cluster(seq)
    assert list(got) == list(exp) == ['a', 'b', 'c']
    for k in got:
        assert np.array_equal(got[k], exp[k])


def test_vg_abs_bound_is_an_upper_bound():
    h = np.linspace(0, 1e5, 2001)
    for vg in ['0.1 Nug(0.0) + 0.9 Sph(20000)', '0.3 Nug(0.0) + 0.7 Hol(25000)',
               '0.1 Nug(0.0) + 0.002 Pow(0.5)', '2.5 Rng(1.0)', '1.0 Exp(20000) + 0.5 Lin(1000)']:
        vals = orc.get_theo_vg_vals(vg, h)
        assert np.nanmax(np.abs(vals)) <= vgs.vg_abs_bound(vg, 1e5) * (1 + 1e-12), vg


def test_shard_bounds():
    b = sdist.shard_bounds(10000, 8)
    assert b[0] == 0 and b[-1] == 10000 and np.all(np.diff(b) == 1250)
    b = sdist.shard_bounds(10, 4)
    assert np.array_equal(b, np.linspace(0, 10, 5, dtype=np.int64))    # ret_mp_idxs, misc.py:601
    b = sdist.shard_bounds(3, 8)                                      # more ranks than steps
    assert b[0] == 0 and b[-1] == 3 and np.all(np.diff(b) >= 0)
    w = np.r_[np.full(100, 9.0), np.full(900, 1.0)]
    b = sdist.shard_bounds(1000, 2, w)
    assert b[0] == 0 and b[-1] == 1000 and abs(w[:b[1]].sum() - w[b[1]:].sum()) <= 9.0


def test_output_file_layout(tmp_path):
    """interp/prepare.py:290-432: dimensions, coordinate variables, per-label
    variables and sett_* attributes."""
    x = np.linspace(500, 9500, 10)
    y = np.linspace(7500, 500, 8)
    tr = pd.date_range('2000-01-01', periods=5, freq='D')
    tv = ncwriter.time_numbers(tr, 'days since 1900-01-01', 'gregorian', 'D')
    assert tv[0] == 36524 and np.all(np.diff(tv) == 1)
    assert np.array_equal(ncwriter.time_numbers(pd.date_range('2000-01-01', periods=3, freq='2D'),
                                                'days since 2000-01-01', 'gregorian', '2D'),
                          [0, 1, 2])   # divided by the numeric prefix of the frequency
    args = [('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)]
    p = ncwriter.create(tmp_path / 'out.nc', x, y, tv, args, np.float32, 'mm', 'precip',
                        'days since 1900-01-01', 'gregorian', 1, {'sett_cell_size': 1000.0})
    h = ncwriter.open_for_update(p)
    fld = np.arange(16, dtype=np.float32).reshape(2, 8)
    h.write('OK', 2, 3, 5, np.tile(fld[:, :1], (1, 10)))
    h.sync()
    h.close()
    from scipy.io import netcdf_file
    if ncwriter.have_netcdf4():
        pytest.skip('netCDF4 backend: layout checked by the netCDF4 library itself')
    f = netcdf_file(str(p), 'r', mmap=False)
    assert f.dimensions == {'dimx': 10, 'dimy': 8, 'dimt': 5}
    assert np.array_equal(f.variables['X'][:], x) and np.array_equal(f.variables['Y'][:], y)
    assert f.variables['Y'][0] > f.variables['Y'][-1]              # Y descending
    assert f.variables['OK'].dimensions == ('dimt', 'dimy', 'dimx')
    assert f.variables['OK'].standard_name == b'precip (OK)'
    assert f.variables['IDW_000'].standard_name == b'precip (IDW_exp_2.0)'
    assert f.variables['time'].units == b'days since 1900-01-01'
    assert f.sett_cell_size == b'1000.0'
    ok = f.variables['OK'][:]
    assert np.isnan(ok[0]).all() and np.array_equal(ok[2, 3:5, 0], [0, 8])
    f.close()


def _frames(n_stn=6, T=4):
    rng = np.random.default_rng(0)
    idx = pd.date_range('2000-01-01', periods=T, freq='D')
    labs = [f'S{i}' for i in range(n_stn)]
    data = pd.DataFrame(rng.gamma(1, 5, (T, n_stn)), index=idx, columns=labs)
    crds = pd.DataFrame({'X': rng.uniform(0, 1e4, n_stn), 'Y': rng.uniform(0, 1e4, n_stn)},
                        index=labs)
    return data, crds


def test_main_configuration_surface(tmp_path):
    """Setter validation and call-order flags behave like interp/data.py."""
    m = SpInterpMain(False)
    with pytest.raises(AssertionError):
        m.interpolate()                                   # verify() first (main.py:76)
    with pytest.raises(AssertionError):
        m.verify()                                        # set_data first (data.py:749)
    data, crds = _frames()
    with pytest.raises(AssertionError):
        m.set_data(data.values, crds)                     # not a DataFrame
    m.set_data(data, crds)
    m.set_out_dir(tmp_path / 'o')
    with pytest.raises(AssertionError):
        m.set_netcdf4_parameters('o.nc', 'mm', 'p', 'days since 1900-01-01', 'gregorian', -1, 1)
    m.set_netcdf4_parameters('o.nc', 'mm', 'p', 'days since 1900-01-01', 'gregorian', 2, 1)
    with pytest.raises(AssertionError):
        m.set_interp_time_parameters('2000-01-05', '2000-01-01', 'D', '%Y-%m-%d')
    m.set_interp_time_parameters('2000-01-01', '2000-01-04', 'D', '%Y-%m-%d')
    with pytest.raises(AssertionError):
        m.set_neighbor_selection_method('bogus')
    with pytest.raises(AssertionError):
        m.set_neighbor_selection_method('nrst')           # n_neighbors missing
    m.set_neighbor_selection_method('all')
    with pytest.raises(AssertionError):
        m.set_misc_settings(min_vg_val=1)                 # must be a float (data.py:725)
    with pytest.raises(AssertionError):
        m.set_misc_settings(min_cutoff_value=5.0, max_cutoff_value=1.0)
    m = SpInterpMain(False)
    m.set_data(data, crds)
    m.set_out_dir(tmp_path / 'o')
    m.set_netcdf4_parameters('o.nc', 'mm', 'p', 'days since 1900-01-01', 'gregorian', 2, 1)
    m.set_interp_time_parameters('2000-01-01', '2000-01-04', 'D', '%Y-%m-%d')
    m.set_neighbor_selection_method('all')
    m.set_misc_settings(cell_size=1000.0)
    with pytest.raises(AssertionError):
        m.verify()                                        # no interpolation method turned on
    with pytest.raises(AssertionError):
        m.turn_ordinary_kriging_est_var_on()              # needs OK first (main.py:440)
    m.turn_inverse_distance_weighting_on([1, 2.5])
    m.turn_nearest_neighbor_on()
    m.verify()
    assert [a[2] for a in m._interp_args] == ['IDW_000', 'IDW_001', 'NNB']
    assert m._interp_crds_orig_shape[0] * m._interp_crds_orig_shape[1] == m._interp_x_crds_msh.size
    assert m._nc_y_crds[0] > m._nc_y_crds[-1]
    assert (tmp_path / 'o' / 'o.nc').exists()
