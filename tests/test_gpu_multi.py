"""SpInterpMain.interpolate() on 2 GPUs against 1 GPU, bit for bit (SURVEY.md section 4
item 4 / section 8e): time-sharded tasks and grid-row-sharded tasks, slabs streamed to the
writer rank over NCCL.  Needs two visible GPUs (skipped otherwise; run with
`gpurun --gpus 2`)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(out_dir, world, row_chunks, port):
    script = str(ROOT / 'scripts' / 'run_main_dist.py')
    env = dict(os.environ)
    if world == 1:
        cmd = [sys.executable, script, str(out_dir), str(row_chunks)]
        for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
            env.pop(k, None)
    else:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               f'--nproc-per-node={world}', '--master-addr', '127.0.0.1', '--master-port',
               str(port), script, str(out_dir), str(row_chunks)]
    r = subprocess.run(cmd, cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert 'DIST MAIN OK' in r.stdout
    return r.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('row_chunks', [0, 3])
def test_two_gpus_equal_one_gpu_bit_for_bit(tmp_path, row_chunks):
    one = tmp_path / 'w1'
    two = tmp_path / 'w2'
    one.mkdir()
    two.mkdir()
    _run(one, 1, row_chunks, 0)
    out = _run(two, 2, row_chunks, 29511 + row_chunks)
    assert "'bytes_received'" in out
    for lab in ('OK', 'IDW_000'):
        a = np.load(one / f'field_{lab}.npy')
        b = np.load(two / f'field_{lab}.npy')
        assert a.tobytes() == b.tobytes(), lab
    s1 = (one / 'stats.csv').read_text()
    s2 = (two / 'stats.csv').read_text()
    assert s1 == s2


def test_one_gpu_row_chunks_equal_whole_grid(tmp_path):
    """Grid-row chunks on ONE GPU (the path a grid too large for HBM takes): same file,
    same statistics as the single-chunk run."""
    a_dir = tmp_path / 'a'
    b_dir = tmp_path / 'b'
    a_dir.mkdir()
    b_dir.mkdir()
    _run(a_dir, 1, 0, 0)
    _run(b_dir, 1, 4, 0)
    for lab in ('OK', 'IDW_000'):
        assert np.load(a_dir / f'field_{lab}.npy').tobytes() == \
            np.load(b_dir / f'field_{lab}.npy').tobytes()
    import pandas as pd
    sa = pd.read_csv(a_dir / 'stats.csv', sep=';', index_col=0)
    sb = pd.read_csv(b_dir / 'stats.csv', sep=';', index_col=0)
    assert list(sa.columns) == list(sb.columns)
    assert np.allclose(sa.values, sb.values, rtol=2e-6, atol=1e-6, equal_nan=True)
