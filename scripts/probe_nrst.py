"""The reference's canonical neighbour setting (test/test_interp.py:101-103: nrst, 50
neighbours) at 1,000 stations on the 1000 x 1000 grid: per-entry-point times of one chunk."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spinterps_b200.engine import ChunkEngine
from tests.synth import VG_C1, make_problem

n_stn = int(os.environ.get('NRST_STN', 1000)); T = int(os.environ.get('NRST_T', 64))
k = int(os.environ.get('NRST_K', 50)); miss = float(os.environ.get('NRST_MISS', 0.0))
p = make_problem(41, n_stn, T, 1000, 1000, miss=0.0)
if miss > 0:      # a few availability groups: some stations missing for blocks of steps
    rng = np.random.default_rng(42)
    for b in range(0, T, 16):
        p['data'][b:b + 16, rng.choice(n_stn, int(miss * n_stn), replace=False)] = np.nan
eng = ChunkEngine()
kw = dict(interp_args=[('OK', None, 'OK')], vgs=[VG_C1] * T, neb_sel_mthd='nrst', n_nebs=k,
          intrp_dtype=np.float32, **p)
eng.submit_chunk(**kw).result(to_host=False)
torch.cuda.synchronize()
eng.trace = []; eng.trace_launches = True
t0 = time.perf_counter()
eng.submit_chunk(**kw).result(to_host=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tr = {k_: round(v['ms'], 3) for k_, v in eng.trace_summary().items()}
print(json.dumps(dict(stations=n_stn, steps=T, neighbours=k, cells=1000000, seconds=dt,
                      cell_steps_per_s=T * 1e6 / dt, entry_points_ms=tr,
                      nrst_systems=eng.stats.get('nrst_systems'))))
