"""Host timeline of a whole config-2 job on one GPU (8 chunks of 1250 steps, different data
per chunk, pageable input): where the time of a short job goes."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spinterps_b200.engine import ChunkEngine
from tests.synth import VG_C1, make_problem

T, chunk = 10000, 1250
p = make_problem(2, 500, T, 1000, 1000, miss=0.2)
base = {k: v for k, v in p.items() if k != 'data'}
kw = dict(interp_args=[('OK', None, 'OK')], intrp_dtype=np.float32, **base)
eng = ChunkEngine()
now = lambda: 1e3 * time.perf_counter()
for rep in range(3):
    torch.cuda.synchronize()
    t00 = now()
    log = []
    pend = None
    for i in range(T // chunk):
        t0 = now()
        nxt = eng.submit_chunk(p['data'][i * chunk:(i + 1) * chunk], vgs=[VG_C1] * chunk, **kw)
        t1 = now()
        if pend is not None:
            f, _ = pend.result(to_host=False)
            del f
        t2 = now()
        log.append((t0 - t00, t1 - t0, t2 - t1))
        pend = nxt
    f, _ = pend.result(to_host=False)
    torch.cuda.synchronize()
    print('rep %d: %.2f ms for the job' % (rep, now() - t00), 'host ms', eng.stats.get('fast_host_ms'))
    print('   ', ' | '.join('%.2f: submit %.2f result %.2f' % l for l in log))
    del f, pend
