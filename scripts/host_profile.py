"""cProfile of the host side of one OK chunk (run on the GPU box)."""
import cProfile, pstats, sys, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from tests.synth import make_problem, VG_C1
from spinterps_b200.engine import ChunkEngine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
p = make_problem(2, 500, T, 1000, 1000, miss=0.2)
eng = ChunkEngine()
args = [('OK', None, 'OK')]
for _ in range(2):
    f, _ = eng.interp_chunk(interp_args=args, vgs=[VG_C1] * T, return_device=True, **p); del f
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
f, _ = eng.interp_chunk(interp_args=args, vgs=[VG_C1] * T, return_device=True, **p)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
