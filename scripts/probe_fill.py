"""HBM roofline of the drop-in fill kernels (C-ABI group 2, device pointers)."""
import sys, json, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from spinterps_b200 import _lib
lib = _lib.load()
PEAK = 6547.8
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
G, N = 1_000_000, 500
rng = np.random.default_rng(0)
cx = torch.from_numpy(rng.uniform(0, 1e6, G)).cuda(); cy = torch.from_numpy(rng.uniform(0, 1e6, G)).cuda()
sx = torch.from_numpy(rng.uniform(0, 1e6, N)).cuda(); sy = torch.from_numpy(rng.uniform(0, 1e6, N)).cuda()
d = torch.empty((G, N), dtype=torch.float64, device='cuda'); v = torch.empty_like(d)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {}
ms = timeit(lambda: _lib.check(lib.spx_fill_dists_2d_mat_dev(cx.data_ptr(), cy.data_ptr(), G, sx.data_ptr(), sy.data_ptr(), N, d.data_ptr(), st)))
res['fill_dists_2d_mat [1e6 x 500] write 4 GB'] = dict(ms=ms, gbs=G * N * 8 / ms / 1e6, frac=G * N * 8 / ms / 1e6 / PEAK)
terms = _lib.parse_vg_str('0.1 Nug(0.0) + 0.9 Sph(20000)')
types = (C.c_int32 * 2)(*[t[0] for t in terms]); sills = (C.c_double * 2)(*[t[1] for t in terms]); ranges = (C.c_double * 2)(*[t[2] for t in terms])
ms = timeit(lambda: _lib.check(lib.spx_fill_vg_var_arr_dev(d.data_ptr(), v.data_ptr(), G, N, 0, 0, 2, types, sills, ranges, 0.0, st)))
res['fill_vg_var_arr Nug+Sph [1e6 x 500] read+write 8 GB'] = dict(ms=ms, gbs=2 * G * N * 8 / ms / 1e6, frac=2 * G * N * 8 / ms / 1e6 / PEAK)
ri = torch.from_numpy(np.sort(rng.choice(G, 400_000, replace=False))).cuda(); ci = torch.from_numpy(np.sort(rng.choice(N, 400, replace=False))).cuda()
sub = torch.empty((400_000, 401), dtype=torch.float64, device='cuda')
ms = timeit(lambda: _lib.check(lib.spx_copy_2d_arr_at_idxs_dev(d.data_ptr(), N, ri.data_ptr(), 400_000, ci.data_ptr(), 400, sub.data_ptr(), 401, st)))
res['copy_2d_arr_at_idxs 4e5 x 400 of [1e6 x 500]'] = dict(ms=ms, gbs=2 * 400_000 * 400 * 8 / ms / 1e6, frac=2 * 400_000 * 400 * 8 / ms / 1e6 / PEAK)
a = torch.empty(G * N, dtype=torch.float64, device='cuda')
ms = timeit(lambda: a.copy_(d.view(-1)))
res['torch copy 4 GB (reference point)'] = dict(ms=ms, gbs=2 * G * N * 8 / ms / 1e6, frac=2 * G * N * 8 / ms / 1e6 / PEAK)
for k, val in res.items(): print(k, {kk: round(vv, 3) for kk, vv in val.items()})
json.dump(res, open(Path(__file__).resolve().parent.parent / 'gpurun_out' / 'probe_fill.json', 'w'), indent=1)
