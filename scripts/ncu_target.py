"""A few chunks of the bench workload through the public call with the fused output stage + delta download:
the target of the `ncu --set full` captures (every hot kernel appears once per chunk)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spinterps_b200.engine import ChunkEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
chunks = [bench.make_chunk(0, v) for v in range(2)]
eng = ChunkEngine()
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * bench.CHUNK_STEPS, intrp_dtype=np.float32,
          round_decimals=bench.NMRL_PRCN, field_stats=True)
for i in range(n):
    out, _ = eng.submit_chunk(**kw, **chunks[i % 2]).result(to_host='packed')
    print(i, out['OK'].nbytes, eng.stats.get('native_submits'), flush=True)
    out['OK'].release()
