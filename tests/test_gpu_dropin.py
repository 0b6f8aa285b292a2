"""GPU parity of the drop-in cyth functions (C-ABI group 1) against the oracle
and the golden known-answer vectors."""
import numpy as np
import pytest

from oracle import spinterp_oracle as orc
from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cy():
    from spinterps_b200 import cyth
    return cyth


@pytest.fixture(scope='module')
def kats():
    return np.load(GOLDEN / 'kats.npz', allow_pickle=False)


def test_fill_dists_bit_exact(cy, kats):
    d = np.full((7, 5), np.nan)
    cy.fill_dists_2d_mat(kats['d_x1'], kats['d_y1'], kats['d_x2'], kats['d_y2'], d)
    assert np.array_equal(d, kats['d_out'])          # bit-exact vs the reference
    rng = np.random.default_rng(3)
    for n1, n2 in [(1, 1), (3, 1001), (257, 64), (1000, 513)]:
        x1, y1 = rng.uniform(0, 1e6, n1), rng.uniform(0, 1e6, n1)
        x2, y2 = rng.uniform(0, 1e6, n2), rng.uniform(0, 1e6, n2)
        got = np.full((n1, n2), np.nan)
        cy.fill_dists_2d_mat(x1, y1, x2, y2, got)
        exp = np.full((n1, n2), np.nan)
        orc.fill_dists_2d_mat(x1, y1, x2, y2, exp)
        assert np.array_equal(got, exp)
    # empty input is a no-op
    cy.fill_dists_2d_mat(np.zeros(0), np.zeros(0), kats['d_x2'], kats['d_y2'], np.zeros((0, 5)))


def test_fill_vg_var_arr(cy, kats):
    d = kats['d_out']
    dd = np.full((7, 7), np.nan)
    cy.fill_dists_2d_mat(kats['d_x1'], kats['d_y1'], kats['d_x1'], kats['d_y1'], dd)
    for vi, vg in enumerate(kats['vg_list']):
        for cov in (0, 1):
            for mv in (0.0, 0.3):
                a = np.full_like(d, np.nan)
                cy.fill_vg_var_arr(d, a, cov, 0, str(vg), mv)
                np.testing.assert_allclose(a, kats[f'vg{vi}_c{cov}_m{int(mv > 0)}_rect'],
                                           rtol=1e-13, atol=1e-15)
                b = np.full_like(dd, np.nan)
                cy.fill_vg_var_arr(dd, b, cov, 1, str(vg), mv)
                np.testing.assert_allclose(b, kats[f'vg{vi}_c{cov}_m{int(mv > 0)}_diag'],
                                           rtol=1e-13, atol=1e-15)


def test_malformed_vg_raises(cy):
    from spinterps_b200._lib import SpxError
    d = np.ones((2, 2))
    with pytest.raises(SpxError):
        cy.fill_vg_var_arr(d, d.copy(), 0, 0, '0.1 Foo(3)', 0.0)
    with pytest.raises(SpxError):
        cy.fill_vg_var_arr(d, d.copy(), 0, 0, 'garbage', 0.0)


def test_gather_theo_idw(cy, kats):
    sub = np.full((4, 6), np.nan)
    cy.copy_2d_arr_at_idxs(kats['cp_arr'], kats['cp_ri'], kats['cp_ci'], sub)
    assert np.array_equal(sub, kats['cp_out'], equal_nan=True)
    np.testing.assert_allclose(
        cy.get_theo_vg_vals('100 Sph(10000) + 10 Exp(1000000)', kats['theo_h']),
        kats['theo_out'], rtol=1e-14)
    w = np.full(3, np.nan)
    s = cy.fill_wts_and_sum(np.array([0.2, 0.5, 1.0]), w, 2.0)
    np.testing.assert_allclose(w, kats['idw_w'], rtol=1e-15)
    np.testing.assert_allclose(s, kats['idw_sum'], rtol=1e-15)
    assert cy.get_mults_sum(kats['idw_w'].copy(), np.array([1., 2., 4.])) == kats['idw_ms']
    dist = np.full(5, np.nan)
    cy.fill_dists_one_pt(3.0, 4.0, kats['d_x2'], kats['d_y2'], dist)
    exp = np.full(5, np.nan)
    orc.fill_dists_one_pt(3.0, 4.0, kats['d_x2'], kats['d_y2'], exp)
    assert np.array_equal(dist, exp)


def test_pie_helper_and_nd_dists(cy, kats):
    """sel_equidist_refs (sector index, members per sector, rank inside the sector) and
    get_nd_dists against the known answers of the compiled reference: bit-exact."""
    rx, ry = kats['pie_rx'], kats['pie_ry']
    n = rx.size
    for n_pies in (3, 4, 8):
        for pi, (px, py) in enumerate(kats['pie_pts']):
            dists, tem = np.zeros(n), np.zeros(n)
            sel = np.zeros(n, dtype=np.int64)
            pidx = np.zeros(n, dtype=np.uint64)
            cts = np.zeros(n_pies, dtype=np.uint64)
            cy.sel_equidist_refs(px, py, rx, ry, n_pies, -1.0, -1, dists, tem, sel, pidx, cts)
            assert np.array_equal(sel, kats[f'pie{n_pies}_sel'][pi])
            assert np.array_equal(pidx.astype(np.int64), kats[f'pie{n_pies}_pidx'][pi])
            assert np.array_equal(cts.astype(np.int64), kats[f'pie{n_pies}_cts'][pi])
            assert np.array_equal(dists, kats[f'pie{n_pies}_dists'][pi])
    # a reference point within the threshold: only the nearest one is selected (rank 0)
    dists, tem = np.zeros(n), np.zeros(n)
    sel = np.zeros(n, dtype=np.int64)
    pidx, cts = np.zeros(n, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    cy.sel_equidist_refs(rx[7] + 1.0, ry[7], rx, ry, 4, 50.0, -1, dists, tem, sel, pidx, cts)
    exp = np.full(n, -1)
    exp[7] = 0
    assert np.array_equal(sel, exp)
    assert np.array_equal(cy.get_nd_dists(kats['nd_pts']), kats['nd_out'])
    with pytest.raises(ValueError):
        cy.sel_equidist_refs(0.0, 0.0, rx, ry, 4, -1.0, -1, dists, tem, sel,
                             np.zeros(n, dtype=np.uint32), cts)
