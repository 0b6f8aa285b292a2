"""Host-side breakdown of the native submit (spx_fast_submit) on the bench workload."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spinterps_b200.engine import ChunkEngine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
print('rank', rank, 'cores in affinity mask', len(os.sched_getaffinity(0)), flush=True)
chunks = [bench.make_chunk(rank, v) for v in range(4)]
eng = ChunkEngine()
eng.solve_stream = os.environ.get('PROBE_SOLVE_STREAM') == '1'
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * bench.CHUNK_STEPS, intrp_dtype=np.float32)
pend = None
for i in range(6):
    eng.submit_chunk(**kw, **chunks[i % 4]).result(to_host=False)
torch.cuda.synchronize()
for rep in range(2):
    acc = np.zeros(6)
    t_sub = t_res = 0.0
    pend = None
    t0 = time.perf_counter()
    for i in range(steps):
        a = time.perf_counter()
        nxt = eng.submit_chunk(**kw, **chunks[i % 4])
        b = time.perf_counter()
        acc += np.array(eng.stats.get('fast_host_ms', [0] * 6))
        if pend is not None:
            pend.result(to_host=False)
        c = time.perf_counter()
        t_sub += b - a
        t_res += c - b
        pend = nxt
    pend.result(to_host=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('rep %d: %.3f ms/step; submit %.3f, result %.3f ms; native phases [wait, scan, plan, upload, solve-q, est-q] = %s'
          % (rep, 1e3 * dt / steps, 1e3 * t_sub / steps, 1e3 * t_res / steps,
             np.round(acc / steps, 3).tolist()), flush=True)
# solve-only and estimate-only timings of one slot
job = next(iter(eng._fast_jobs.values()))
import ctypes as C
from spinterps_b200 import _lib
e = C.c_float(); s = C.c_float()
_lib.check(eng.lib.spx_fast_times(job['handle'], 0, C.byref(e), C.byref(s)))
print('slot 0: estimate %.3f ms, solve phase %.3f ms (overlapped)' % (e.value, s.value))
# timeline of the last 4 chunks (ring slots), relative to the oldest one's solve start
torch.cuda.synchronize()
n_slots = job['cfg'].n_slots
order = [(job['next_slot'] + k) % n_slots for k in range(n_slots)]     # oldest first
tl = (C.c_float * 4)()
for k in order:
    _lib.check(eng.lib.spx_fast_timeline(job['handle'], order[0], k, tl))
    print('slot %d: solve [%.3f, %.3f]  estimate [%.3f, %.3f] ms' % (k, tl[0], tl[1], tl[2], tl[3]))
if os.environ.get('PROBE_CPROFILE') == '1':
    import cProfile, pstats
    pr = cProfile.Profile()
    pend = None
    pr.enable()
    for i in range(20):
        nxt = eng.submit_chunk(**kw, **chunks[i % 4])
        if pend is not None:
            pend.result(to_host=False)
        pend = nxt
    pend.result(to_host=False)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
