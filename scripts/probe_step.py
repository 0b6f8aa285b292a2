"""Per-entry-point breakdown of one resident bench step (+ optional phase clocks of
k_downdate_reg when the library exports spx_debug_dd_clocks)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spinterps_b200.engine import ChunkEngine  # noqa: E402
from spinterps_b200 import _lib  # noqa: E402

steps = int(os.environ.get('PROBE_STEPS', bench.CHUNK_STEPS))
bench.CHUNK_STEPS = steps
chunks = [bench.make_chunk(0, v) for v in range(3)]
eng = ChunkEngine()
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * steps, intrp_dtype=np.float32)
for i in range(3):
    eng.submit_chunk(**kw, **chunks[i % 3]).result(to_host=False)
torch.cuda.synchronize()
lib = _lib.load()
has_clk = hasattr(lib, 'spx_debug_dd_clocks')
buf = (C.c_ulonglong * 8)()
if has_clk:
    lib.spx_debug_dd_clocks(buf, 1)
eng.trace_launches = True
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
n = 4
e0.record()
for i in range(n):
    eng.submit_chunk(**kw, **chunks[i % 3]).result(to_host=False)
e1.record()
torch.cuda.synchronize()
print('ms/step', e0.elapsed_time(e1) / n)
print(json.dumps({k: round(v['ms'] / n, 4) for k, v in eng.trace_summary().items()}))
if has_clk:
    lib.spx_debug_dd_clocks(buf, 0)
    tot = sum(buf)
    print('dd phase clocks (thread 0, summed over CTAs):',
          [round(100.0 * b / max(tot, 1), 1) for b in buf], 'cycles/CTA', tot / (n * steps))

# host side: wall time of submit_chunk alone (device work is asynchronous)
import time, cProfile, pstats, io
eng.trace_launches = False
torch.cuda.synchronize()
ts = []
pend = []
for i in range(6):
    t0 = time.perf_counter()
    pend.append(eng.submit_chunk(**kw, **chunks[i % 3]))
    ts.append((time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter()
for q in pend:
    q.result(to_host=False)
t_res = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
print('submit_chunk host ms:', [round(t, 2) for t in ts], ' result() x6 ms', round(t_res, 2))
pr = cProfile.Profile()
pr.enable()
for i in range(4):
    eng.submit_chunk(**kw, **chunks[i % 3]).result(to_host=False)
pr.disable()
s_ = io.StringIO()
pstats.Stats(pr, stream=s_).sort_stats('cumulative').print_stats(70)
print(s_.getvalue()[:14000])
