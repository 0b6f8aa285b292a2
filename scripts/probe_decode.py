"""Host decode throughput of the 2-byte field transport (spx_unpack_field_host) by thread
count, and the bare device -> pinned host copy rates next to it."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('SPX_HOST_THREADS', '4')
from spinterps_b200 import _lib
lib = _lib.load()
T, G, d = 1250, 1000000, 2
if len(sys.argv) > 1:
    T = int(sys.argv[1])
stride = int(lib.spx_pack_stride(G))
rng = np.random.default_rng(0)
codes = rng.integers(0, 6000, size=(T, stride), dtype=np.uint16)
hdr = np.zeros(T, dtype=_lib.PACK_ROW_DTYPE)
out = np.empty((T, G), dtype=np.float32)
out[:] = 0
print('cores in affinity mask:', len(os.sched_getaffinity(0)))
for nt in (1, 2, 4, 8, 12, 16, 24, 32):
    if nt > len(os.sched_getaffinity(0)):
        break
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(lib.spx_unpack_field_host(hdr.ctypes.data, codes.ctypes.data, T, G, d,
                                             out.ctypes.data, G, nt))
        best = min(best, time.perf_counter() - t0)
    print('decode %2d threads: %7.1f ms  %6.1f GB/s of f32 output' % (nt, best * 1e3, T * G * 4 / best / 1e9), flush=True)
try:
    import torch
    if torch.cuda.is_available():
        n = T * G
        dev = torch.empty(n, dtype=torch.float32, device='cuda')
        pin = torch.empty(n, dtype=torch.float32).pin_memory()
        for frac, lab in ((1.0, 'f32 field'), (0.5, 'u16 codes')):
            m = int(n * frac)
            best = 1e9
            for rep in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                pin[:m].copy_(dev[:m], non_blocking=True); torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            print('D2H %s: %.1f ms  %.1f GB/s' % (lab, best * 1e3, m * 4 / best / 1e9))
except Exception as e:
    print('no torch/cuda:', e)
