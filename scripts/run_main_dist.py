"""SpInterpMain.interpolate() on 1 rank or under torchrun: every rank interpolates its
(time chunk x grid-row chunk) tasks, the writer rank receives the slabs over NCCL and
writes.  Compares the file with the oracle and stores the raw fields as .npy so that runs
with different world sizes can be compared bit for bit.  Usage:
    python scripts/run_main_dist.py <out_dir> [row_chunks]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
        scripts/run_main_dist.py <out_dir> [row_chunks]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, pandas as pd, torch, torch.distributed as dist
from spinterps_b200.main import SpInterpMain
from spinterps_b200 import ncwriter
from oracle import spinterp_oracle as orc

rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
out_dir = Path(sys.argv[1])
row_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
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
if row_chunks:
    m._grid_row_chunks_forced = row_chunks
m.verify(); m.interpolate()
print('rank', rank, 'gather', m.gather_stats, flush=True)
if rank == 0:
    exp, _ = orc.interp_chunk(m._data_df.values, m._crds_df['X'].values, m._crds_df['Y'].values,
                              m._interp_x_crds_msh, m._interp_y_crds_msh, m._interp_crds_orig_shape,
                              m._interp_args, vgs=[vg] * T, intrp_dtype=np.float32)
    h = ncwriter.open_for_read(m._nc_file_path); ny, nx = m._interp_crds_orig_shape
    worst = 0.0
    for lab in ('OK', 'IDW_000'):
        ref = np.round(exp[lab], 2).reshape(T, ny, nx)
        full = np.empty((T, ny, nx), dtype=np.float32)
        for t in range(T):
            got = h.read(lab, t)
            assert not np.isnan(got).any(), (lab, t, int(np.isnan(got).sum()), got.shape,
                                             np.argwhere(np.isnan(got))[:5].tolist())
            full[t] = got
            worst = max(worst, float(np.abs(got - ref[t]).max()))
        np.save(out_dir / f'field_{lab}.npy', full)
    h.close()
    print('DIST MAIN OK world', world, 'row_chunks', row_chunks,
          'max |diff| after 2-decimal rounding', worst)
    assert worst <= 0.0101
if world > 1:
    dist.destroy_process_group()
