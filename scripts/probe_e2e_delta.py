"""Stage times of the end-to-end leg with the delta transport (one chunk of the bench
workload, nothing overlapped), then the pipelined loop with a host-side timeline."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spinterps_b200.engine import ChunkEngine

eng = ChunkEngine()
T = bench.CHUNK_STEPS
chunks = [bench.make_chunk(0, v) for v in range(2)]
for c in chunks:
    c['data'] = torch.from_numpy(c['data']).pin_memory().numpy()
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * T, intrp_dtype=np.float32,
          round_decimals=bench.NMRL_PRCN, field_stats=True)
for i in range(4):
    eng.submit_chunk(**kw, **chunks[i % 2]).result(to_host='packed')[0]['OK'].release()
torch.cuda.synchronize()
dl = eng._dl


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


pend = eng.submit_chunk(**kw, **chunks[0])
for fn in pend.deferred:
    fn()
pend.deferred = []
fld = pend.flds['OK']
st = torch.empty((5, T), dtype=torch.float64, device='cuda')
torch.cuda.synchronize()
for name, a in (('round+stats+encode', dict(round_here=True, stats=st, write_back=False)),
                ('round+encode', dict(round_here=True, stats=None, write_back=False)),
                ('round+stats+encode+write-back', dict(round_here=True, stats=st, write_back=True)),
                ('verify+encode (rounded input)', dict())):
    for rep in range(2):
        e0 = ev()
        t0 = time.perf_counter()
        tk = dl.start(fld, bench.NMRL_PRCN, **a)
        e1 = ev()
        pf = dl.wait(tk)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print('%-32s kernel %.2f ms; start->landed %.2f ms for %.1f MB (%.3f B/cell)' % (
            name, e0.elapsed_time(e1), 1e3 * (t1 - t0), pf.nbytes / 1e6,
            pf.nbytes / (T * bench.NY * bench.NX)), flush=True)
        dl.release(tk)

import concurrent.futures
pool = concurrent.futures.ThreadPoolExecutor(1, initializer=lambda: torch.cuda.set_device(eng.device))


def now():
    return 1e3 * time.perf_counter()


for rep in range(2):
    torch.cuda.synchronize()
    t00 = now()
    log = []
    futs = [None] * 2
    pend = None

    def land(pend, h, k):
        t0 = now()
        out, _ = pend.finish_packed(h)
        log.append(('landed%d' % k, now() - t00, now() - t0))
        out['OK'].release()

    def drain(pend, k):
        t0 = now()
        if futs[k % 2] is not None:
            futs[k % 2].result()
        t1 = now()
        h = pend.start_packed()
        futs[k % 2] = pool.submit(land, pend, h, k)
        log.append(('drain%d' % k, t0 - t00, t1 - t0, now() - t1))
    n = 8
    for k in range(n):
        t0 = now()
        nxt = eng.submit_chunk(**kw, **chunks[k % 2])
        log.append(('submit%d' % k, t0 - t00, now() - t0))
        if pend is not None:
            drain(pend, k - 1)
        pend = nxt
    drain(pend, n - 1)
    for f in futs:
        if f is not None:
            f.result()
    torch.cuda.synchronize()
    print('pipelined rep %d: %.2f ms per chunk' % (rep, (now() - t00) / n))
    if rep == 1:
        for l in sorted(log, key=lambda l: l[1]):
            print('  ', l[0], ' '.join('%.2f' % v for v in l[1:]))
