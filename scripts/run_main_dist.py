"""SpInterpMain under torchrun: every rank interpolates its time block, rank 0
writes; compares the file with a single-rank oracle run.  Usage:
torchrun --nproc-per-node N scripts/run_main_dist.py <out_dir>"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, pandas as pd, torch, torch.distributed as dist
from spinterps_b200.main import SpInterpMain
from spinterps_b200 import ncwriter
from oracle import spinterp_oracle as orc

rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
out_dir = Path(sys.argv[1])
rng = np.random.default_rng(4)
n_stn, T = 40, 23
idx = pd.date_range('2001-03-01', periods=T, freq='D')
labs = [f'P{i:03d}' for i in range(n_stn)]
vals = rng.gamma(1.0, 5.0, (T, n_stn)); vals[rng.random((T, n_stn)) < 0.15] = np.nan
data = pd.DataFrame(vals, index=idx, columns=labs)
crds = pd.DataFrame({'X': rng.uniform(0, 6e4, n_stn), 'Y': rng.uniform(0, 5e4, n_stn)}, index=labs)
vg = '0.1 Nug(0.0) + 0.9 Sph(20000)'
m = SpInterpMain(False)
m.set_data(data, crds); m.set_vgs_ser(pd.Series([vg] * T, index=idx, dtype=object))
m.set_out_dir(out_dir)
m.set_netcdf4_parameters('p.nc', 'mm', 'precip', 'days since 1900-01-01', 'gregorian', 2, 1)
m.set_interp_time_parameters('2001-03-01', '2001-03-23', 'D', '%Y-%m-%d')
m.set_neighbor_selection_method('all'); m.set_misc_settings(cell_size=1500.0, max_steps_per_chunk=5)
m.turn_ordinary_kriging_on(); m.turn_inverse_distance_weighting_on([2])
m.verify(); m.interpolate()
if rank == 0:
    exp, _ = orc.interp_chunk(m._data_df.values, m._crds_df['X'].values, m._crds_df['Y'].values,
                              m._interp_x_crds_msh, m._interp_y_crds_msh, m._interp_crds_orig_shape,
                              m._interp_args, vgs=[vg] * T, intrp_dtype=np.float32)
    h = ncwriter.open_for_read(m._nc_file_path); ny, nx = m._interp_crds_orig_shape
    worst = 0.0
    for lab in ('OK', 'IDW_000'):
        ref = np.round(exp[lab], 2).reshape(T, ny, nx)
        for t in range(T):
            got = h.read(lab, t); assert not np.isnan(got).any(), (lab, t)
            worst = max(worst, float(np.abs(got - ref[t]).max()))
    h.close()
    print('DIST MAIN OK world', dist.get_world_size(), 'max |diff| after 2-decimal rounding', worst)
    assert worst <= 0.0101
dist.destroy_process_group()
