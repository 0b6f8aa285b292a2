"""GPU timeline of the resident bench loop (torch.profiler / CUPTI): kernel start / end
times of a few steady-state chunks, to see where the GPU idles between kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spinterps_b200.engine import ChunkEngine
from torch.profiler import profile, ProfilerActivity

chunks = [bench.make_chunk(0, v) for v in range(4)]
eng = ChunkEngine()
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * bench.CHUNK_STEPS, intrp_dtype=np.float32)

def run(n):
    pend = None
    for i in range(n):
        nxt = eng.submit_chunk(**kw, **chunks[i % 4])
        if pend is not None:
            pend.result(to_host=False)
        pend = nxt
    pend.result(to_host=False)

run(6)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(8)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = None
busy = 0.0
for e in evs:
    s, en = e.time_range.start - t0, e.time_range.end - t0
    gap = (s - prev_end) if prev_end is not None else 0.0
    busy += en - s
    print('%9.1f us  +%7.1f gap  %8.1f us  %s' % (s, gap, en - s, e.name[:60]))
    prev_end = max(prev_end or 0, en)
print('span %.1f us busy %.1f us' % (prev_end, busy))
