"""Time only the estimate contraction for the bench shape under the current env knobs."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spinterps_b200.engine import ChunkEngine
eng = ChunkEngine(); p = bench.make_chunk(0); T = bench.CHUNK_STEPS
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * T, intrp_dtype=np.float32)
eng.profile_gemm = True
for _ in range(4):
    eng.gemm_events = []
    f, _ = eng.interp_chunk(return_device=True, **kw, **p); torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in eng.gemm_events]
print({k: os.environ.get(k) for k in ('SPX_GEMM_NT', 'SPX_GEMM_STAGES', 'SPX_GEMM_WARPS')}, 'gemm ms', ms, 'TF', eng.stats['gemm_flop'] / ms[0] / 1e9)
