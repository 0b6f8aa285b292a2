"""Per-line wall time of the host side of submit_chunk (sys.settrace on the engine's
planning functions); run on the GPU box."""
import os, sys, time, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spinterps_b200 import engine as E

chunks = [bench.make_chunk(0, v) for v in range(3)]
eng = E.ChunkEngine()
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * bench.CHUNK_STEPS, intrp_dtype=np.float32)
for i in range(3):
    eng.submit_chunk(**kw, **chunks[i % 3]).result(to_host=False)
torch.cuda.synchronize()
names = {'_interp_chunk', '_krige', '_solve_downdate', '_local_plan', '_local_neighbours',
         'availability_groups', '_dev', '_dev_pack', '_mask_lists', '_fetch_async', 'deferred',
         'result', '_output_stage', '_arena_take'}
codes = {}
acc = collections.defaultdict(float)
cnt = collections.defaultdict(int)
state = {}

def local(frame, event, arg):
    now = time.perf_counter()
    fid = id(frame)
    if fid in state:
        key, t0 = state[fid]
        acc[key] += now - t0
        cnt[key] += 1
    if event == 'return':
        state.pop(fid, None)
        return local
    state[fid] = ((frame.f_code.co_name, frame.f_lineno), time.perf_counter())
    return local

def glob(frame, event, arg):
    if event == 'call' and frame.f_code.co_name in names and 'engine.py' in frame.f_code.co_filename:
        return local
    return None

N = 6
sys.settrace(glob)
for i in range(N):
    eng.submit_chunk(**kw, **chunks[i % 3]).result(to_host=False)
sys.settrace(None)
torch.cuda.synchronize()
src = open(os.path.join(ROOT, 'spinterps_b200', 'engine.py')).read().splitlines()
tot = sum(acc.values())
print('total traced ms/chunk', 1e3 * tot / N)
for (fn, ln), t in sorted(acc.items(), key=lambda kv: -kv[1])[:45]:
    print('%7.3f ms %4d  %-18s %5d  %s' % (1e3 * t / N, cnt[(fn, ln)] // N, fn, ln, src[ln - 1].strip()[:90]))
