import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spinterps_b200.engine import ChunkEngine
eng = ChunkEngine()
p = bench.make_chunk(0)
T = bench.CHUNK_STEPS
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * T, intrp_dtype=np.float32)
pin_out = [torch.empty((T, 1000000), dtype=torch.float32).pin_memory() for _ in range(2)]
copy_stream = torch.cuda.Stream(); copy_done = [None, None]
def now(): return 1e3 * time.perf_counter()
for rep in range(2):
    torch.cuda.synchronize(); t00 = now(); log = []
    pend = None
    def drain(pend, k):
        t0 = now(); flds, _ = pend.result(to_host=False); t1 = now()
        if copy_done[k % 2] is not None: copy_done[k % 2].synchronize()
        t2 = now()
        copy_stream.wait_event(pend.done_event)
        with torch.cuda.stream(copy_stream):
            pin_out[k % 2].copy_(flds['OK'], non_blocking=True)
            flds['OK'].record_stream(copy_stream)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        copy_done[k % 2] = ev
        t3 = now(); log.append(('drain%d' % k, t0 - t00, t1 - t0, t2 - t1, t3 - t2))
    for k in range(5):
        t0 = now(); nxt = eng.submit_chunk(**kw, **p); log.append(('submit%d' % k, t0 - t00, now() - t0))
        if pend is not None: drain(pend, k - 1)
        pend = nxt
    drain(pend, 4); t0 = now(); copy_stream.synchronize(); torch.cuda.synchronize(); log.append(('final', t0 - t00, now() - t0))
    print('rep', rep, 'total %.1f' % (now() - t00))
    for l in log: print('  ', l[0], ' '.join('%.1f' % v for v in l[1:]))
