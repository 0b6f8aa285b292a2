"""Pin the NumPy oracle against the unmodified reference (golden fixtures) and
the known-answer vectors of SURVEY.md section 8c.  CPU only."""
import numpy as np
import pytest

from oracle import spinterp_oracle as orc
from tests.golden_util import CASES, GOLDEN, load_case, rel_err


@pytest.fixture(scope='module')
def kats():
    return np.load(GOLDEN / 'kats.npz', allow_pickle=False)


def test_survey_known_answers():
    # SURVEY.md 8c
    h = np.linspace(0, 1e6, 10)
    exp = [0, 102.8346868942621, 104.86582880967408, 106.32120558828558,
           107.36402861884274, 108.11124397162438, 108.64664716763387,
           109.03028032135595, 109.30516548777199, 109.50212931632136]
    np.testing.assert_allclose(
        orc.get_theo_vg_vals('100 Sph(10000) + 10 Exp(1000000)', h), exp, rtol=1e-15)
    d = np.full((2, 2), np.nan)
    x, y = np.array([0., 3.]), np.array([0., 4.])
    orc.fill_dists_2d_mat(x, y, x, y, d)
    assert np.array_equal(d, [[0, 5], [5, 0]])
    v = np.full((2, 2), np.nan)
    orc.fill_vg_var_arr(d, v, 0, 1, '0.1 Nug(0.0) + 0.9 Sph(20000)', 0.0)
    np.testing.assert_allclose(v, [[0.1, 0.1003375], [0.1003375, 0.1]], rtol=1e-7)  # printed to 7 digits in the survey
    w = np.full(3, np.nan)
    s = orc.fill_wts_and_sum(np.array([0.2, 0.5, 1.0]), w, 2.0)
    assert s == 29.999999999999996 and w[0] == 24.999999999999996
    assert orc.get_mults_sum(w, np.array([1., 2., 4.])) / s == 1.2333333333333334


def test_free_function_kats(kats):
    d = np.full((7, 5), np.nan)
    orc.fill_dists_2d_mat(kats['d_x1'], kats['d_y1'], kats['d_x2'], kats['d_y2'], d)
    # numpy's (..)**0.5 is sqrt; libm pow(x, 0.5) is correctly rounded too
    assert np.array_equal(d, kats['d_out'])
    dd = np.full((7, 7), np.nan)
    orc.fill_dists_2d_mat(kats['d_x1'], kats['d_y1'], kats['d_x1'], kats['d_y1'], dd)
    for vi, vg in enumerate(kats['vg_list']):
        for cov in (0, 1):
            for mv in (0.0, 0.3):
                a = np.full_like(d, np.nan)
                orc.fill_vg_var_arr(d, a, cov, 0, str(vg), mv)
                np.testing.assert_allclose(
                    a, kats[f'vg{vi}_c{cov}_m{int(mv > 0)}_rect'], rtol=2e-15, atol=1e-300)
                b = np.full_like(dd, np.nan)
                orc.fill_vg_var_arr(dd, b, cov, 1, str(vg), mv)
                np.testing.assert_allclose(
                    b, kats[f'vg{vi}_c{cov}_m{int(mv > 0)}_diag'], rtol=2e-15, atol=1e-300)
    np.testing.assert_allclose(
        orc.get_theo_vg_vals('100 Sph(10000) + 10 Exp(1000000)', kats['theo_h']),
        kats['theo_out'], rtol=1e-15)
    w = np.full(3, np.nan)
    assert orc.fill_wts_and_sum(np.array([0.2, 0.5, 1.0]), w, 2.0) == kats['idw_sum']
    assert np.array_equal(w, kats['idw_w'])
    assert orc.get_mults_sum(w, np.array([1., 2., 4.])) == kats['idw_ms']
    sub = np.full((4, 6), np.nan)
    orc.copy_2d_arr_at_idxs(kats['cp_arr'], kats['cp_ri'], kats['cp_ci'], sub)
    assert np.array_equal(sub, kats['cp_out'], equal_nan=True)
    for key in kats.files:
        if key.startswith('nugget__'):
            assert orc.check_full_nuggetness(key[8:], 1e-4) == bool(kats[key])
    # pie helper and pairwise distances: index outputs bit-exact
    rx, ry = kats['pie_rx'], kats['pie_ry']
    for n_pies in (3, 4, 8):
        for pi, (px, py) in enumerate(kats['pie_pts']):
            n = rx.size
            dists, tem = np.zeros(n), np.zeros(n)
            sel, pidx = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
            cts = np.zeros(n_pies, dtype=np.int64)
            orc.sel_equidist_refs(px, py, rx, ry, n_pies, -1.0, -1, dists, tem, sel, pidx, cts)
            assert np.array_equal(sel, kats[f'pie{n_pies}_sel'][pi])
            assert np.array_equal(pidx, kats[f'pie{n_pies}_pidx'][pi])
            assert np.array_equal(cts, kats[f'pie{n_pies}_cts'][pi])
            assert np.array_equal(dists, kats[f'pie{n_pies}_dists'][pi])
    assert np.array_equal(orc.get_nd_dists(kats['nd_pts']), kats['nd_out'])


@pytest.mark.parametrize('faithful', [True, False])
@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference(name, faithful):
    case, outs = load_case(name)
    flds, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=faithful, **case)
    assert set(flds) == set(outs)
    for lab, ref in outs.items():
        # same algorithm, same LAPACK: only summation-order noise is allowed
        tol = 1e-12 if lab.startswith(('IDW', 'NNB')) else 1e-9
        if name == 'f_vg_families':
            tol = 1e-6  # ill-conditioned Gau/Pow systems amplify gemv-vs-gemm noise
        assert rel_err(flds[lab], ref) <= tol, (name, lab, rel_err(flds[lab], ref))


@pytest.mark.skipif(not __import__('pathlib').Path('/root/reference/interp/steps.py').exists(),
                    reason='needs the reference sources (build container only)')
def test_fixtures_regenerate_from_the_reference(tmp_path):
    """The committed fixtures ARE outputs of the unmodified reference: the generating script,
    run again in a fresh process (pyximport build of the reference's Cython module), writes
    byte-identical files."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(GOLDEN / 'make_golden.py'), '--out', str(tmp_path)],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    names = sorted(p.name for p in tmp_path.glob('*.npz'))
    assert set(names) >= {f'{c}.npz' for c in CASES} | {'kats.npz'}
    for n in names:
        assert (tmp_path / n).read_bytes() == (GOLDEN / n).read_bytes(), n
