"""CPU restatements of two kernel algorithms whose claim is "identical results, different
schedule".  They run without a GPU and pin the LOGIC of the schedules; the kernels
themselves are compared on the GPU (tests/test_gpu_engine.py:
test_topk_warp_kernel_equals_the_insertion_kernel,
test_nrst_solve_thread_and_warp_substitutions_agree).

* ``k_topk_warp`` (csrc/spx_nrst.cu): the neighbour row of a cell is the set of the k
  smallest (distance, station index) pairs.  The kernel finds it without sorting: bisection
  on the 64-bit pattern of the distances with a count per round, early exit when a
  threshold separates exactly k keys, ties at the threshold resolved by index.
* ``k_nrst_solve`` thread path: forward / backward substitution in dot-product form, four
  rows at a time, must give every row its updates in the order of the column sweeps
  (IEEE fma is not associative: a different order would change the last bits).
"""
from fractions import Fraction

import numpy as np
import pytest

U64_MAX = 2 ** 64 - 1


def select_by_bisection(keys, valid, k):
    """The selection of k_topk_warp, statement by statement."""
    keys = np.where(valid, keys, np.uint64(U64_MAX)).astype(np.uint64)
    nv = int(valid.sum())
    kk = min(k, nv)
    thr, n_eq, rounds = 0, 0, 0
    if kk == nv:
        thr = U64_MAX
    elif kk > 0:
        lo, hi, exact = int(keys[valid].min()), int(keys[valid].max()), False
        while lo < hi:              # invariant: count(<= lo - 1) < kk <= count(<= hi)
            mid = lo + ((hi - lo) >> 1)
            cnt = int((keys <= np.uint64(mid)).sum())
            rounds += 1
            if cnt == kk:
                thr, exact = mid + 1, True
                break
            if cnt > kk:
                hi = mid
            else:
                lo = mid + 1
        if not exact:
            thr = lo
            below = 0 if lo == 0 else int((keys <= np.uint64(lo - 1)).sum())
            n_eq = kk - below
    out, eq_seen = [], 0
    for s in range(keys.size):
        key = int(keys[s])
        eq = key == thr and thr != U64_MAX
        if key < thr or (eq and eq_seen < n_eq):
            out.append(s)
        eq_seen += int(eq)
    return out + [-1] * (k - len(out)), rounds


@pytest.mark.parametrize('mode', ['random', 'many_ties', 'coarse'])
def test_bisection_selection_is_the_k_smallest_by_distance_then_index(mode):
    rng = np.random.default_rng({'random': 0, 'many_ties': 1, 'coarse': 2}[mode])
    worst_rounds = 0
    for trial in range(400):
        n = int(rng.integers(1, 200))
        k = int(rng.integers(1, 70))
        if mode == 'random':
            d = rng.random(n) * 1e5
        elif mode == 'many_ties':
            d = rng.integers(0, 5, n).astype(np.float64)
        else:
            d = np.round(rng.random(n), 1)
        valid = rng.random(n) < 0.8 if trial % 2 else np.ones(n, bool)
        got, rounds = select_by_bisection(d.view(np.uint64), valid, k)
        worst_rounds = max(worst_rounds, rounds)
        cand = np.flatnonzero(valid)
        sel = np.sort(cand[np.lexsort((cand, d[cand]))][:k])
        assert got == sel.tolist() + [-1] * (k - sel.size), (mode, trial)
    assert worst_rounds <= 64            # one round per key bit at most


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))     # exactly rounded once


def _column_sweeps(S, y):
    """Warp path of k_nrst_solve: axpy per column, unit lower L and upper U in one matrix."""
    m = y.size
    y = y.copy()
    for c in range(m - 1):
        for i in range(c + 1, m):
            y[i] = _fma(-S[i, c], y[c], y[i])
    for c in range(m - 1, -1, -1):
        y[c] = y[c] / S[c, c]
        for i in range(c):
            y[i] = _fma(-S[i, c], y[c], y[i])
    return y


def _dot_product_rows(S, y):
    """Thread path: four rows at a time in registers, no stores inside the inner loops."""
    m = y.size
    y = y.copy()
    for i0 in range(1, m, 4):
        rows = [min(i0 + r, m - 1) for r in range(4)]
        a = [y[i] for i in rows]
        for c in range(i0):
            a = [_fma(-S[i, c], y[c], ar) for i, ar in zip(rows, a)]
        y[i0] = a[0]
        for r in range(1, 4):
            if i0 + r < m:
                for q in range(r):
                    a[r] = _fma(-S[i0 + r, i0 + q], a[q], a[r])
                y[i0 + r] = a[r]
    for i0 in range(m - 1, -1, -4):
        rows = [max(i0 - r, 0) for r in range(4)]
        a = [y[i] for i in rows]
        for c in range(m - 1, i0, -1):
            a = [_fma(-S[i, c], y[c], ar) for i, ar in zip(rows, a)]
        x = [a[0] / S[i0, i0]]
        y[i0] = x[0]
        for r in range(1, 4):
            if i0 - r < 0:
                break
            for q in range(r):
                a[r] = _fma(-S[i0 - r, i0 - q], x[q], a[r])
            x.append(a[r] / S[i0 - r, i0 - r])
            y[i0 - r] = x[r]
    return y


@pytest.mark.parametrize('m', [1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 16, 17])
def test_dot_product_substitution_keeps_the_order_of_the_column_sweeps(m):
    rng = np.random.default_rng(m)
    for _ in range(3):
        S = rng.normal(size=(m, m))
        y = rng.normal(size=m)
        assert np.array_equal(_column_sweeps(S, y), _dot_product_rows(S, y))
