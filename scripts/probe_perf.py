"""Quick device-side timing probe of the engine on a C2-shaped chunk (not the bench)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from tests.synth import make_problem, VG_C1
from spinterps_b200.engine import ChunkEngine

def dgemm_peak(n=8192, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * n ** 3 / best / 1e9

def main():
    n_stn = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ny = nx = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    miss = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
    print('cuBLAS DGEMM 8192^3 TFLOP/s:', dgemm_peak())
    p = make_problem(2, n_stn, T, ny, nx, miss=miss)
    eng = ChunkEngine()
    import os
    eng.sync_timing = bool(int(os.environ.get('SPX_SYNC_TIMING', '0')))
    for args in ([('OK', None, 'OK')], [('IDW', None, 'IDW_000', 2.0)], [('NNB', None, 'NNB')]):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            flds, _ = eng.interp_chunk(interp_args=args, vgs=[VG_C1] * T, return_device=True, **p)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); wall = time.time() - t0
            cs = T * ny * nx
            print(args[0][0], 'rep', rep, f'dev {ms:.1f} ms wall {wall*1e3:.1f} ms  cell-steps/s {cs/ms*1e3:.3e}',
                  'stats', eng.stats, 'TFLOP/s(gemm)', eng.stats.get('gemm_flop', 0) / ms / 1e9, 'timing', {k: round(v, 2) for k, v in eng.timing.items()})
            del flds
main()
