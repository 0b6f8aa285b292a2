"""Stand-alone kriging classes (cyth/interpmthds.pyx:251-765) on the GPU against
the reference's own outputs (tests/golden/kats.npz, written by make_golden.py)."""
import numpy as np
import pytest

from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-9, atol=1e-10)


@pytest.fixture(scope='module')
def k():
    return np.load(GOLDEN / 'kats.npz', allow_pickle=False)


def test_survey_known_answer():
    from spinterps_b200.kriging import OrdinaryKriging
    c = OrdinaryKriging(np.array([0., 10., 0.]), np.array([0., 0., 10.]), np.array([1., 2., 4.]),
                        np.array([5., 2.]), np.array([5., 1.]), '0.1 Nug(0.0) + 0.9 Sph(20)')
    c.krige()
    np.testing.assert_allclose(c.zk, [2.452383392372142, 1.5809023584822606], rtol=1e-12)
    np.testing.assert_allclose(c.mus, [0.03545941744630898, 0.01605501917316882], rtol=1e-10)
    np.testing.assert_allclose(c.est_vars, [0.5928691165263567, 0.39322759196103607], rtol=1e-12)
    np.testing.assert_allclose(c.lambdas[0], [0.27380830381392807, 0.3630958480930356,
                                              0.3630958480930357], rtol=1e-12)


def test_classes_match_reference(k):
    from spinterps_b200 import kriging as kr
    xi, yi, zi, xk, yk = k['kc_xi'], k['kc_yi'], k['kc_zi'], k['kc_xk'], k['kc_yk']
    si, sk, model = k['kc_si'], k['kc_sk'], str(k['kc_model'])
    c = kr.OrdinaryKriging(xi, yi, zi, xk, yk, model)
    c.krige()
    for a in ('zk', 'lambdas', 'mus', 'est_vars', 'rhss', 'in_vars'):
        np.testing.assert_allclose(getattr(c, a), k['ok_' + a], err_msg=a, **TOL)
    assert c.in_vars[0, 0] == 0.0 and c.rhss[0, 3] == 0.0      # zero diagonal / zero distance
    c = kr.SimpleKriging(xi, yi, zi, xk, yk, model)
    c.krige()
    for a in ('zk', 'lambdas', 'est_covars', 'rhss', 'in_covars'):
        np.testing.assert_allclose(getattr(c, a), k['sk_' + a], err_msg=a, **TOL)
    c = kr.ExternalDriftKriging(xi, yi, zi, si[0], xk, yk, sk[0], model)
    c.krige()
    for a in ('zk', 'lambdas', 'mus_1', 'mus_2'):
        np.testing.assert_allclose(getattr(c, a), k['edk_' + a], err_msg=a, rtol=1e-8, atol=1e-9)
    c = kr.ExternalDriftKriging_MD(xi, yi, zi, si, xk, yk, sk, model)
    c.krige()
    for a in ('zk', 'lambdas', 'mus_arr'):
        np.testing.assert_allclose(getattr(c, a), k['md_' + a], err_msg=a, rtol=1e-8, atol=1e-9)
    c = kr.OrdinaryIndicatorKriging(xi, yi, zi, xk, yk, 4.0, model)
    c.ikrige()
    np.testing.assert_allclose(c.ik, k['oik_ik'], **TOL)
    np.testing.assert_allclose(c.est_vars, k['oik_est_vars'], **TOL)
    c = kr.SimpleIndicatorKriging(xi, yi, zi, xk, yk, 4.0, model)
    c.ikrige()
    np.testing.assert_allclose(c.ik, k['sik_ik'], **TOL)
    np.testing.assert_allclose(c.est_covars, k['sik_est_covars'], **TOL)
