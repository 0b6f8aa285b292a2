"""Time the estimate contraction under the current env knobs (SPX_GEMM_DIRECT / _NT / _STAGES /
_WARPS): the bench shape through the dense path (kpad 504) or, with 'c4', one IDW exponent of
config 4 (kpad 2008, 1000 steps x 1e6 cells)."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spinterps_b200.engine import ChunkEngine
from tests.synth import CONFIG_CHUNK, config_problem

which = sys.argv[1] if len(sys.argv) > 1 else 'dense'
eng = ChunkEngine()
if which == 'c4':
    p, kw = config_problem('C4', CONFIG_CHUNK['C4'], seed_shift=1)
    kw = dict(kw, interp_args=kw['interp_args'][:1])
    base = dict(p)
else:
    base = bench.make_chunk(0)
    kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * bench.CHUNK_STEPS)
    eng.local_support = False
eng.profile_gemm = True
res = []
for _ in range(3):
    eng.kernel_events = []
    eng.submit_chunk(intrp_dtype=np.float32, **kw, **base).result(to_host=False)
    torch.cuda.synchronize()
    eng.collect_profile()
    ev = [e for e in eng.kernel_events if e[0] == 'k_estimate_gemm']
    ms = [(e[3] if e[4] is None else e[3].elapsed_time(e[4])) for e in ev]
    res.append((sum(ms), sum(e[2] for e in ev)))
ms, flop = res[-1]
print(which, {k: os.environ.get(k) for k in ('SPX_GEMM_DIRECT', 'SPX_GEMM_NT', 'SPX_GEMM_STAGES',
                                               'SPX_GEMM_WARPS')},
      'gemm ms %.2f  TFLOP/s %.2f' % (ms, flop / ms / 1e9), [round(r[0], 2) for r in res])
