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


def elev(x, y):
    """Smooth analytic elevation surface: the external drift of config 3."""
    return 400 + 0.0006 * x + 0.0003 * y + 120 * np.sin(x / 2.3e5) * np.cos(y / 1.7e5)


# steps one rank processes per bench step (a time chunk), per configuration
CONFIG_CHUNK = {'C1': 365, 'C2': 1250, 'C3': 250, 'C4': 1000, 'C5': 100}


def config_problem(cfg, n_steps=None, seed_shift=0):
    """The synthetic inputs of BASELINE.json's five configurations:
    returns (problem dict, engine keyword arguments).  n_steps: how many time steps to
    generate (default: the configuration's own); seed_shift: different time steps for
    different ranks / chunks of one job (stations and grid stay the job's)."""
    def steps(full):
        return int(full if n_steps is None else n_steps)

    def fresh_data(p, seed, n_stn, miss=0.0):
        if seed_shift:
            rng = np.random.default_rng(seed * 1000 + seed_shift)
            T = p['data'].shape[0]
            d = rng.gamma(1.0, 5.0, size=(T, n_stn))
            if miss > 0:
                d[rng.random((T, n_stn)) < miss] = np.nan
            p['data'] = d
        return p

    if cfg == 'C1':
        T = steps(365)
        p = fresh_data(make_problem(1, 100, T, 200, 200), 1, 100)
        return p, dict(interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)],
                       vgs=[VG_C1] * T)
    if cfg == 'C2':
        T = steps(10000)
        p = fresh_data(make_problem(2, 500, T, 1000, 1000, miss=0.2), 2, 500, 0.2)
        return p, dict(interp_args=[('OK', None, 'OK')], vgs=[VG_C1] * T)
    if cfg == 'C3':
        T = steps(5000)
        p = fresh_data(make_problem(3, 300, T, 2000, 2000), 3, 300)
        rng = np.random.default_rng(33 + seed_shift)
        vgs = ['%0.5f Nug(0.0) + %0.5f Sph(%0.5f)' % (rng.uniform(0, 0.2), rng.uniform(0.5, 1.5),
                                                        rng.uniform(1e4, 5e4)) for _ in range(T)]
        kw = dict(interp_args=[('EDK', None, 'EDK')], vgs=vgs,
                  drft_arrs=elev(p['cell_xs'], p['cell_ys'])[None, :],
                  stns_drft=elev(p['stn_xs'], p['stn_ys'])[:, None])
        return p, kw
    if cfg == 'C4':
        T = steps(20000)
        p = fresh_data(make_problem(4, 2000, T, 1000, 1000), 4, 2000)
        args = [('IDW', None, 'IDW_%03d' % i, float(e)) for i, e in enumerate((1, 2, 3, 5))]
        return p, dict(interp_args=args)
    if cfg == 'C5':
        T = steps(2000)
        p = fresh_data(make_problem(5, 1000, T, 4000, 4000), 5, 1000)
        cx, cy = p['cell_xs'], p['cell_ys']
        mask = ((cx - 2.0e6) / 1.8e6) ** 2 + ((cy - 2.0e6) / 1.4e6) ** 2 <= 1.0   # ~49 %
        p['cell_xs'], p['cell_ys'] = cx[mask], cy[mask]
        kw = dict(interp_args=[('OK', None, 'OK'), ('SK', None, 'SK')], vgs=[VG_C1] * T,
                  cntn_idxs=mask)
        return p, kw
    raise ValueError(f'unknown configuration {cfg}')
