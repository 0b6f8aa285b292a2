"""cProfile of the FIRST chunk of a job in a fresh process (one-time set-up costs)."""
import cProfile, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spinterps_b200.engine import ChunkEngine
from tests.synth import VG_C1, make_problem
T = 1250
p = make_problem(2, 500, 2 * T, 1000, 1000, miss=0.2)
base = {k: v for k, v in p.items() if k != 'data'}
kw = dict(interp_args=[('OK', None, 'OK')], intrp_dtype=np.float32, vgs=[VG_C1] * T, **base)
torch.zeros(1, device='cuda'); torch.cuda.synchronize()
eng = ChunkEngine()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
f, _ = eng.submit_chunk(p['data'][:T], **kw).result(to_host=False)
torch.cuda.synchronize()
pr.disable()
print('first chunk %.1f ms' % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter()
f2, _ = eng.submit_chunk(p['data'][T:], **kw).result(to_host=False)
torch.cuda.synchronize()
print('second chunk %.1f ms' % (1e3 * (time.perf_counter() - t0)))
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
