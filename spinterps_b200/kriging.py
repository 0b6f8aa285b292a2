"""Stand-alone kriging classes of the reference's native module
(cyth/interpmthds.pyx:251-765): ``OrdinaryKriging``, ``SimpleKriging``,
``ExternalDriftKriging``, ``ExternalDriftKriging_MD`` and the two indicator
variants -- thin wrappers over the same sm_100a kernels as the gridded path
(distance fill, variogram fill, LU factor / solve through the C-ABI).

Same constructor arguments, ``krige()`` / ``ikrige()`` and result attributes.
Their semantics differ from the ``SpInterpMain`` path and are kept as they are in
the reference: the diagonal of the system is 0 (only h > g pairs are evaluated,
pyx:322-329), a right-hand-side entry at zero distance is skipped (pyx:344-345),
there is no ``min_vg_val`` cut, and ``SimpleKriging`` builds its matrix from a
covariance that is still 0.0 at construction time (pyx:397-399 vs :413), i.e.
off-diagonal ``-gamma`` and a zero diagonal.  ``np.linalg.pinv`` (pyx:335) is
replaced by LU: results agree for non-singular systems.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Device:
    """Shared plumbing: device matrices from coordinates, LU inverse."""

    def __init__(self):
        _lib.require_gpu()
        self.lib = _lib.load()
        self.dev = torch.device('cuda', torch.cuda.current_device())

    def up(self, a):
        return torch.from_numpy(_f64(a)).to(self.dev)

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def dists(self, x1, y1, x2, y2):
        d = torch.empty((x1.numel(), x2.numel()), dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.spx_fill_dists_2d_mat_dev(
            x1.data_ptr(), y1.data_ptr(), x1.numel(), x2.data_ptr(), y2.data_ptr(), x2.numel(),
            d.data_ptr(), self.stream()), 'fill_dists_2d_mat')
        return d

    def vg(self, dists, terms):
        """Sum of nested variogram terms, no covariance flip, no cut."""
        out = torch.empty_like(dists)
        n = len(terms)
        types = (C.c_int32 * n)(*[t[0] for t in terms])
        sills = (C.c_double * n)(*[t[1] for t in terms])
        ranges = (C.c_double * n)(*[t[2] for t in terms])
        _lib.check(self.lib.spx_fill_vg_var_arr_dev(
            dists.data_ptr(), out.data_ptr(), dists.shape[0], dists.shape[1], 0, 0, n, types,
            sills, ranges, float('-inf'), self.stream()), 'fill_vg_var_arr')
        return out

    def inverse(self, A):
        """A^-1 of a symmetric device matrix through the batched LU kernels."""
        m = A.shape[0]
        work = A.contiguous().clone()
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=self.dev)   # noqa: E731
        i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=self.dev)   # noqa: E731
        t = [i32([m]), i32([_lib.KRG_KINDS['SK']]), i32([0]), i64([0]), i64([0]), i64([0]),
             torch.arange(m, dtype=torch.int32, device=self.dev)]
        piv = torch.empty(m, dtype=torch.int32, device=self.dev)
        info = torch.zeros(1, dtype=torch.int32, device=self.dev)
        S = _lib.spx_systems()
        S.n_sys, S.n_drifts = 1, 0
        (S.sys_n, S.sys_kind, S.sys_vg, S.sys_stn_off, S.sys_w_off, S.sys_piv_off,
         S.stn_list) = (x.data_ptr() for x in t)
        S.work, S.piv, S.info, S.max_m = work.data_ptr(), piv.data_ptr(), info.data_ptr(), m
        _lib.check(self.lib.spx_krige_factor_dev(C.byref(S), self.stream()), 'factor')
        inv = torch.empty((m, m), dtype=torch.float64, device=self.dev)
        rt = [torch.zeros(m, dtype=torch.int32, device=self.dev),
              torch.full((m,), 2, dtype=torch.int32, device=self.dev),
              torch.arange(m, dtype=torch.int32, device=self.dev),
              torch.full((m,), -1, dtype=torch.int64, device=self.dev)]
        R = _lib.spx_rhs()
        R.n_rhs = m
        R.rhs_sys, R.rhs_kind, R.rhs_arg, R.rhs_row = (x.data_ptr() for x in rt)
        R.data, R.n_stn, R.kpad = inv.data_ptr(), m, 8
        R.coef, R.resid = inv.data_ptr(), None
        R.dense, R.dense_ld = inv.data_ptr(), m
        _lib.check(self.lib.spx_krige_solve_dev(C.byref(S), C.byref(R), self.stream()), 'solve')
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError(
                f'singular kriging system (zero pivot at column {int(info.item())})')
        return inv


class _Base:
    _n_extra = 0

    def _setup(self, xi, yi, zi, xk, yk, model):
        self.xi, self.yi, self.zi = _f64(xi), _f64(yi), _f64(zi)
        self.xk, self.yk = _f64(xk), _f64(yk)
        self.model = bytes(model, 'utf-8')
        self._model_str = model
        self.in_count = self.xi.shape[0]
        self.out_count = self.xk.shape[0]
        self.zk = np.zeros(self.out_count)
        self.lambdas = np.zeros((self.out_count, self.in_count))

    def _terms(self):
        terms = _lib.parse_vg_str(self._model_str)      # range clamp 1e-5, pyx:303
        self.sills = [t[1] for t in terms]
        self.ranges = [t[2] for t in terms]
        self.vgs = [_lib.VG_NAMES[t[0]].encode() for t in terms]
        return terms

    def _station_block(self, dv, terms):
        xi, yi = dv.up(self.xi), dv.up(self.yi)
        d_in = dv.dists(xi, yi, xi, yi)
        g_in = dv.vg(d_in, terms)
        g_in.fill_diagonal_(0.0)                          # only h != g pairs are evaluated
        d_out = dv.dists(dv.up(self.xk), dv.up(self.yk), xi, yi)
        g_out = torch.where(d_out == 0.0, torch.zeros_like(d_out), dv.vg(d_out, terms))
        self.in_dists = d_in.cpu().numpy()
        return g_in, g_out, d_out


class OrdinaryKriging(_Base):
    """cyth/interpmthds.pyx:251-362."""

    def __init__(self, xi, yi, zi, xk, yk, model='1.0 Sph(2)'):
        self._setup(xi, yi, zi, xk, yk, model)
        self.mus = np.zeros(self.out_count)
        self.est_vars = np.zeros(self.out_count)

    def _border(self, dv, n, no):
        return torch.ones((n, 1), dtype=torch.float64, device=dv.dev), \
            torch.ones((no, 1), dtype=torch.float64, device=dv.dev)

    def krige(self):
        dv = _Device()
        terms = self._terms()
        n, no = self.in_count, self.out_count
        g_in, g_out, _ = self._station_block(dv, terms)
        b_in, b_out = self._border(dv, n, no)
        nb = b_in.shape[1]
        A = torch.zeros((n + nb, n + nb), dtype=torch.float64, device=dv.dev)
        A[:n, :n] = g_in
        A[:n, n:] = b_in
        A[n:, :n] = b_in.T
        rhs = torch.cat([g_out, b_out], dim=1)
        inv = dv.inverse(A)
        lam = rhs @ inv.T                                  # pyx:352 for every target
        z = dv.up(self.zi)
        self.in_vars = A.cpu().numpy()
        self.in_vars_inv = inv.cpu().numpy()
        self.rhss = rhs.cpu().numpy()
        self.lambdas = lam[:, :n].cpu().numpy()
        self.zk = (lam[:, :n] @ z).cpu().numpy()
        self._finish(lam, rhs, n)

    def _finish(self, lam, rhs, n):
        self.mus = lam[:, n].cpu().numpy()
        ev = (lam[:, :n] * rhs[:, :n]).sum(dim=1) + lam[:, n]
        self.est_vars = torch.clamp(ev, min=0.0).cpu().numpy()       # pyx:358-359


class ExternalDriftKriging_MD(OrdinaryKriging):
    """cyth/interpmthds.pyx:588-719 (si [n_drifts, n_in], sk [n_drifts, n_out])."""

    def __init__(self, xi, yi, zi, si, xk, yk, sk, model='1.0 Sph(2)'):
        xi = np.asarray(xi)
        si, sk = np.asarray(si), np.asarray(sk)
        assert len(xi.shape) == 1
        assert len(si.shape) == 2
        assert xi.shape[0] == np.asarray(yi).shape[0] == np.asarray(zi).shape[0] == si.shape[1], (
            'Observation points and drift shapes are unequal!')
        assert np.asarray(xk).shape[0] == np.asarray(yk).shape[0] == sk.shape[1], (
            'Resulting points and drifts shapes are unequal!')
        assert si.shape[0] == sk.shape[0], 'Observation and reulting drifts have unequal shapes!'
        self._setup(xi, yi, zi, xk, yk, model)
        self.si, self.sk = _f64(si), _f64(sk)
        self.n_drifts = sk.shape[0]
        self.mus_arr = np.zeros((self.n_drifts + 1, self.out_count))

    def _border(self, dv, n, no):
        ones_i = torch.ones((n, 1), dtype=torch.float64, device=dv.dev)
        ones_o = torch.ones((no, 1), dtype=torch.float64, device=dv.dev)
        return torch.cat([ones_i, dv.up(self.si).T], dim=1), \
            torch.cat([ones_o, dv.up(self.sk).T], dim=1)

    def _finish(self, lam, rhs, n):
        self.mus_arr = lam[:, n:].T.contiguous().cpu().numpy()


class ExternalDriftKriging(ExternalDriftKriging_MD):
    """cyth/interpmthds.pyx:474-585 (one drift: si [n_in], sk [n_out])."""

    def __init__(self, xi, yi, zi, si, xk, yk, sk, model='1.0 Sph(2)'):
        ExternalDriftKriging_MD.__init__(self, xi, yi, zi, np.asarray(si)[None, :], xk, yk,
                                         np.asarray(sk)[None, :], model)
        self.si, self.sk = _f64(si), _f64(sk)
        self.mus_1 = np.zeros(self.out_count)
        self.mus_2 = np.zeros(self.out_count)

    def _border(self, dv, n, no):
        ones_i = torch.ones((n, 1), dtype=torch.float64, device=dv.dev)
        ones_o = torch.ones((no, 1), dtype=torch.float64, device=dv.dev)
        return torch.cat([ones_i, dv.up(self.si)[:, None]], dim=1), \
            torch.cat([ones_o, dv.up(self.sk)[:, None]], dim=1)

    def _finish(self, lam, rhs, n):
        self.mus_1 = lam[:, n].cpu().numpy()
        self.mus_2 = lam[:, n + 1].cpu().numpy()


class SimpleKriging(_Base):
    """cyth/interpmthds.pyx:365-471 including its construction-time covariance."""

    def __init__(self, xi, yi, zi, xk, yk, model='1.0 Sph(2)'):
        self._setup(xi, yi, zi, xk, yk, model)
        self.est_covars = np.zeros(self.out_count)
        self.covar = 0.0

    def krige(self):
        dv = _Device()
        terms = self._terms()
        self.covar = float(sum(self.sills))
        g_in, g_out, d_out = self._station_block(dv, terms)
        A = -g_in                                            # 0.0 - gamma, zero diagonal
        rhs = self.covar - g_out                             # zero distance -> covar
        inv = dv.inverse(A)
        lam = rhs @ inv.T
        z = dv.up(self.zi)
        self.in_covars = A.cpu().numpy()
        self.in_covars_inv = inv.cpu().numpy()
        self.out_covars = rhs.cpu().numpy()
        self.rhss = self.out_covars.copy()
        self.lambdas = lam.cpu().numpy()
        self.zk = (lam @ z).cpu().numpy()
        ec = self.covar - (lam * rhs).sum(dim=1)
        self.est_covars = torch.clamp(ec, min=0.0).cpu().numpy()    # pyx:465-466


class OrdinaryIndicatorKriging(OrdinaryKriging):
    """cyth/interpmthds.pyx:722-742."""

    def __init__(self, xi, yi, zi, xk, yk, lim=1, model='1.0 Sph(2)'):
        OrdinaryKriging.__init__(self, xi, yi, zi, xk, yk, model=model)
        self.lim = lim
        self.ixi = np.where(self.zi <= self.lim, 1., 0.)
        self.ik = np.zeros(self.out_count)

    def ikrige(self):
        self.krige()
        self.ik = np.maximum(0.0, (self.lambdas * self.ixi[None, :]).sum(axis=1))
        self.est_vars = np.maximum(0.0, self.ik * (1. - self.ik))


class SimpleIndicatorKriging(SimpleKriging):
    """cyth/interpmthds.pyx:745-765."""

    def __init__(self, xi, yi, zi, xk, yk, lim=1, model='1.0 Sph(2)'):
        SimpleKriging.__init__(self, xi, yi, zi, xk, yk, model=model)
        self.lim = lim
        self.ixi = np.where(self.zi <= self.lim, 1., 0.)
        self.ik = np.zeros(self.out_count)

    def ikrige(self):
        self.krige()
        self.ik = np.maximum(0.0, (self.lambdas * self.ixi[None, :]).sum(axis=1))
        self.est_covars = np.maximum(0.0, self.ik * (1. - self.ik))
