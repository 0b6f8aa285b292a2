"""One chunk of the bench workload, then the fused output stage on its field (ncu target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from spinterps_b200.engine import ChunkEngine
eng = ChunkEngine()
T = bench.CHUNK_STEPS
c = bench.make_chunk(0, 0)
kw = dict(interp_args=bench.INTERP_ARGS, vgs=[bench.VG] * T, intrp_dtype=np.float32,
          round_decimals=bench.NMRL_PRCN, field_stats=True)
for i in range(2):
    out = eng.submit_chunk(**kw, **c).result(to_host='packed')[0]['OK']
    print(i, out.nbytes, flush=True)
    out.release()
