"""The resident (value) loop of bench.py alone, optionally under torch.distributed, to
separate per-rank host cost from everything else bench.py does.
   python scripts/value_loop.py [steps]            (CUDA_VISIBLE_DEVICES picks the GPU)
   torchrun --nproc-per-node N scripts/value_loop.py [steps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spinterps_b200.engine import ChunkEngine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1 and os.environ.get('NO_DIST') != '1':
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
chunks = [bench.make_chunk(rank, v) for v in range(4)]
if os.environ.get('VL_PINNED_IN') == '1':   # station data in pinned host memory
    for c in chunks:
        c['data'] = torch.from_numpy(c['data']).pin_memory().numpy()
eng = ChunkEngine()
if os.environ.get('VL_NATIVE') == '0':
    eng.native_submit = False
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * bench.CHUNK_STEPS, intrp_dtype=np.float32)

DEPTH = int(os.environ.get('VL_DEPTH', '2'))


def run(n):
    import collections
    pend = collections.deque()
    t_sub = 0.0
    for i in range(n):
        if len(pend) == DEPTH:
            pend.popleft().result(to_host=False)
        t0 = time.perf_counter()
        pend.append(eng.submit_chunk(**kw, **chunks[i % 4]))
        t_sub += time.perf_counter() - t0
    while pend:
        pend.popleft().result(to_host=False)
    return t_sub

keep = []
if os.environ.get('VL_PINNED') == '1':      # bench.py's pinned e2e buffers
    keep = [torch.empty((bench.CHUNK_STEPS, bench.NY * bench.NX), dtype=torch.float32).pin_memory()
            for _ in range(2)]
if os.environ.get('VL_SAMPLER') == '1':
    smp = bench.ClockSampler(local)
    smp.start()
if os.environ.get('VL_PROF') == '1':
    eng.profile_gemm = True
if os.environ.get('VL_FORK') == '1':
    pass
run(5)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    ts = run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('rank %d rep %d: %.3f ms/step (submit %.3f ms/step) native_submits=%s affinity %d cores; '
          'native host ms of the last chunk [slot wait, scan, plan, uploads, solve, estimate] %s' % (
              rank, rep, 1e3 * dt / steps, 1e3 * ts / steps, eng.stats.get('native_submits'),
              len(os.sched_getaffinity(0)),
              ['%.3f' % v for v in eng.stats.get('fast_host_ms', [])]), flush=True)

if os.environ.get('VL_PROF') == '1':
    eng.collect_profile()
    est = sorted(e[3] for e in eng.kernel_events[-steps:])
    sol = sorted(eng.solve_ms[-steps:])
    job = next(iter(eng._fast_jobs.values()))
    ptrs = {'d_arena': job['d_arena'].data_ptr()}
    for i, k in enumerate(job['keep']):
        if isinstance(k, dict):
            for name, t in k.items():
                if torch.is_tensor(t):
                    ptrs['%d.%s' % (i, name)] = t.data_ptr()
                elif isinstance(t, (tuple, list)):
                    for j, tt in enumerate(t):
                        if torch.is_tensor(tt):
                            ptrs['%d.%s%d' % (i, name, j)] = tt.data_ptr()
    print('rank %d estimate ms min/med/max %.3f %.3f %.3f  solve ms min/med/max %.3f %.3f %.3f' % (
        rank, est[0], est[len(est) // 2], est[-1], sol[0], sol[len(sol) // 2], sol[-1]))
    print('rank %d ptrs' % rank, {k: hex(v) for k, v in ptrs.items()})
    print('rank %d mem' % rank, torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20)
