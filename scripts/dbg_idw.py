import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spinterps_b200.engine import ChunkEngine
from tests.synth import VG_C1
rng = np.random.default_rng(4)
n_stn, T = 40, 23
vals = rng.gamma(1.0, 5.0, (T, n_stn)); vals[rng.random((T, n_stn)) < 0.15] = np.nan
sx = rng.uniform(0, 6e4, n_stn); sy = rng.uniform(0, 5e4, n_stn)
cs = 1500.0
# same grid as SpInterpMain would build is not needed: any grid
nx, ny = 39, 31
gx = sx.min() + cs * (0.5 + np.arange(nx)); gy = sy.max() - cs * (0.5 + np.arange(ny))
mx, my = np.meshgrid(gx, gy)
kw = dict(stn_xs=sx, stn_ys=sy, cell_xs=mx.ravel(), cell_ys=my.ravel(), grid_shape=(ny, nx),
          interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)], intrp_dtype=np.float32)
bounds = [0, 4, 9, 13, 18, 23]
for mode in ('seq-pipelined', 'seq-sync', 'fresh'):
    eng = ChunkEngine()
    res = []
    if mode == 'seq-pipelined':
        prev = None
        for i in range(5):
            a, b = bounds[i], bounds[i + 1]
            cur = eng.submit_chunk(vals[a:b], vgs=[VG_C1] * (b - a), round_decimals=2, field_stats=True, **kw)
            if prev is not None:
                res.append(prev.result()[0])
            prev = cur
        res.append(prev.result()[0])
    else:
        for i in range(5):
            a, b = bounds[i], bounds[i + 1]
            if mode == 'fresh':
                eng = ChunkEngine()
            res.append(eng.submit_chunk(vals[a:b], vgs=[VG_C1] * (b - a), round_decimals=2, field_stats=True, **kw).result()[0])
    print(mode, [(int(np.isnan(r['OK']).sum()), [int(np.isnan(r['IDW_000'][t]).sum()) for t in range(r['IDW_000'].shape[0])]) for r in res])
