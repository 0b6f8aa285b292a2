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
    eng.kernel_events = []
    f, _ = eng.interp_chunk(return_device=True, **kw, **p); torch.cuda.synchronize()
    ms = [(n, a.elapsed_time(b)) for n, _, _, a, b in eng.kernel_events]
print(os.environ.get('SPX_LOCAL_ROWS'), ms)
