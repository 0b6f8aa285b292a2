"""Run the BASELINE.json configurations (SURVEY.md section 8d inputs) through the
engine on ONE GPU, time-chunked like SpInterpMain.interpolate, and spot-check
parity against the oracle on sampled cells / steps.

    python scripts/run_configs.py C1 C2 ... [--steps-limit N]

Prints one JSON line per configuration (also written to gpurun_out/).
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import spinterp_oracle as orc  # noqa: E402
from spinterps_b200.engine import ChunkEngine  # noqa: E402
from tests.golden_util import rel_err  # noqa: E402
from tests.synth import VG_C1, make_problem  # noqa: E402


def elev(x, y):
    return 400 + 0.0006 * x + 0.0003 * y + 120 * np.sin(x / 2.3e5) * np.cos(y / 1.7e5)


def build(cfg, steps_limit):
    if cfg == 'C1':
        p = make_problem(1, 100, 365, 200, 200)
        return p, dict(interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)],
                       vgs=[VG_C1] * 365), 365
    if cfg == 'C2':
        T = min(10000, steps_limit or 10000)
        p = make_problem(2, 500, T, 1000, 1000, miss=0.2)
        return p, dict(interp_args=[('OK', None, 'OK')], vgs=[VG_C1] * T), 1250
    if cfg == 'C3':
        T = min(5000, steps_limit or 5000)
        p = make_problem(3, 300, T, 2000, 2000)
        rng = np.random.default_rng(33)
        vgs = ['%0.5f Nug(0.0) + %0.5f Sph(%0.5f)' % (rng.uniform(0, 0.2), rng.uniform(0.5, 1.5),
                                                        rng.uniform(1e4, 5e4)) for _ in range(T)]
        kw = dict(interp_args=[('EDK', None, 'EDK')], vgs=vgs,
                  drft_arrs=elev(p['cell_xs'], p['cell_ys'])[None, :],
                  stns_drft=elev(p['stn_xs'], p['stn_ys'])[:, None])
        return p, kw, 250
    if cfg == 'C4':
        T = min(20000, steps_limit or 20000)
        p = make_problem(4, 2000, T, 1000, 1000)
        args = [('IDW', None, 'IDW_%03d' % i, float(e)) for i, e in enumerate((1, 2, 3, 5))]
        return p, dict(interp_args=args), 1000
    if cfg == 'C5':
        T = min(2000, steps_limit or 2000)
        p = make_problem(5, 1000, T, 4000, 4000)
        cx, cy = p['cell_xs'], p['cell_ys']
        mask = ((cx - 2.0e6) / 1.8e6) ** 2 + ((cy - 2.0e6) / 1.4e6) ** 2 <= 1.0   # ~49 %
        p['cell_xs'], p['cell_ys'] = cx[mask], cy[mask]
        kw = dict(interp_args=[('OK', None, 'OK'), ('SK', None, 'SK')], vgs=[VG_C1] * T,
                  cntn_idxs=mask)
        return p, kw, 100
    raise SystemExit(f'unknown config {cfg}')


def spot_check(p, kw, t0, t1, flds_dev, n_cells_chk=1500, n_steps_chk=3, seed=0):
    """Oracle on a random subset of cells and steps of this chunk."""
    rng = np.random.default_rng(seed)
    G = p['cell_xs'].size
    ci = np.sort(rng.choice(G, size=min(n_cells_chk, G), replace=False))
    ti = np.sort(rng.choice(np.arange(t0, t1), size=min(n_steps_chk, t1 - t0), replace=False))
    okw = dict(kw)
    okw.pop('cntn_idxs', None)
    if 'vgs' in okw and okw['vgs'] is not None:
        okw['vgs'] = [kw['vgs'][t] for t in ti]
    if okw.get('drft_arrs') is not None:
        okw['drft_arrs'] = kw['drft_arrs'][:, ci]
    exp, _ = orc.interp_chunk(p['data'][ti], p['stn_xs'], p['stn_ys'], p['cell_xs'][ci],
                              p['cell_ys'][ci], (1, ci.size), intrp_dtype=np.float32,
                              faithful=False, **okw)
    pos = ci
    if kw.get('cntn_idxs') is not None:
        pos = np.where(kw['cntn_idxs'])[0][ci]
    errs = {}
    for lab, ref in exp.items():
        got = flds_dev[lab][torch.as_tensor(ti - t0, device='cuda')][:, torch.as_tensor(pos, device='cuda')]
        got = got.cpu().numpy()
        floor = max(1e-3, 0.01 * float(np.nanmax(np.abs(ref))))
        errs[lab] = rel_err(got, ref, floor)
    return errs


def run(cfg, steps_limit):
    p, kw, chunk = build(cfg, steps_limit)
    eng = ChunkEngine()
    T = p['data'].shape[0]
    bounds = list(range(0, T, chunk)) + [T]
    n_labels = len([a for a in kw['interp_args']])
    G = p['cell_xs'].size
    base = {k: v for k, v in p.items() if k != 'data'}

    def submit(i):
        t0, t1 = bounds[i], bounds[i + 1]
        ckw = dict(kw)
        if ckw.get('vgs') is not None:
            ckw['vgs'] = kw['vgs'][t0:t1]
        return eng.submit_chunk(p['data'][t0:t1], intrp_dtype=np.float32, **base, **ckw)

    # warm-up on the first chunk (also the parity spot check)
    pend = submit(0)
    flds, _ = pend.result(to_host=False)
    errs = spot_check(p, kw, bounds[0], bounds[1], flds)
    del flds, pend
    torch.cuda.synchronize()
    t_beg = time.perf_counter()
    pend = None
    stats = {}
    for i in range(len(bounds) - 1):
        nxt = submit(i)
        if pend is not None:
            f, _ = pend.result(to_host=False)
            del f
        pend = nxt
        for k_, v_ in eng.stats.items():
            if isinstance(v_, (int, float)):
                stats[k_] = stats.get(k_, 0) + v_
    f, _ = pend.result(to_host=False)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_beg
    line = {
        'config': cfg, 'n_gpus': 1, 'stations': int(p['stn_xs'].size), 'steps': int(T),
        'cells': int(G), 'labels': [a[2] for a in kw['interp_args']],
        'chunk_steps': chunk, 'seconds': wall,
        'cell_steps_per_s': n_labels * G * T / wall,
        'gemm_tflops': stats.get('gemm_flop', 0) / wall / 1e12,
        'spot_check_rel_err_f32_store': errs,
        'max_mem_gb': torch.cuda.max_memory_allocated() / 1e9,
        'stats': {k_: int(v_) for k_, v_ in stats.items()},
    }
    print(json.dumps(line), flush=True)
    out = ROOT / 'gpurun_out'
    out.mkdir(exist_ok=True)
    with open(out / f'config_{cfg}.json', 'w') as fh:
        json.dump(line, fh)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='+')
    ap.add_argument('--steps-limit', type=int, default=0)
    a = ap.parse_args()
    for c in a.configs:
        run(c, a.steps_limit)
        torch.cuda.empty_cache()
