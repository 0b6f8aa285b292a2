"""Seeded synthetic inputs of the BASELINE.json config shapes."""
import numpy as np

VG_C1 = '0.1 Nug(0.0) + 0.9 Sph(20000)'


def make_problem(seed, n_stn, n_steps, ny, nx, cell=1000.0, miss=0.0, min_sep=1.0):
    """Stations uniform in the grid's bounding box (min separation enforced),
    gamma(1, 5) data, i.i.d. Bernoulli(miss) NaNs, cell-centre grid with Y
    descending (interp/prepare.py:193-210)."""
    rng = np.random.default_rng(seed)
    side_x, side_y = nx * cell, ny * cell
    xs = rng.uniform(0, side_x, n_stn)
    ys = rng.uniform(0, side_y, n_stn)
    for _ in range(20):  # re-draw the rare too-close pairs
        d = np.hypot(xs[:, None] - xs[None, :], ys[:, None] - ys[None, :])
        np.fill_diagonal(d, np.inf)
        bad = np.where(d.min(axis=1) < min_sep)[0]
        if not bad.size:
            break
        xs[bad] = rng.uniform(0, side_x, bad.size)
        ys[bad] = rng.uniform(0, side_y, bad.size)
    data = rng.gamma(1.0, 5.0, size=(n_steps, n_stn))
    if miss > 0:
        data[rng.random((n_steps, n_stn)) < miss] = np.nan
    gx = 0.5 * cell + cell * np.arange(nx)
    gy = side_y - 0.5 * cell - cell * np.arange(ny)
    mx, my = np.meshgrid(gx, gy)
    return dict(data=data, stn_xs=xs, stn_ys=ys, cell_xs=mx.ravel(), cell_ys=my.ravel(),
                grid_shape=(ny, nx))
