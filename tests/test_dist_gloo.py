"""The N>1 path on CPU: two gloo ranks shard the time axis, compute their slabs
and gather them on the writer rank; the result equals the single-process one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spinterps_b200 import dist as sdist


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _field(beg, end, ncell):
    t = torch.arange(beg, end, dtype=torch.float32)[:, None]
    c = torch.arange(ncell, dtype=torch.float32)[None, :]
    return {'OK': t * 1000 + c, 'IDW_000': -(t + c)}


def _worker(rank, world, port, n_steps, ncell, weights, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        arrs = {'x': np.arange(5.0) + 1, 'm': np.eye(3, dtype=np.float32)} if rank == 0 else {}
        got = sdist.broadcast_inputs(arrs, src=0)
        assert np.array_equal(got['x'], np.arange(5.0) + 1) and got['m'].dtype == np.float32
        beg, end = sdist.my_shard(n_steps, weights)
        out = sdist.run_time_sharded(lambda b, e: _field(b, e, ncell), n_steps, weights, dst=0)
        if rank == 0:
            q.put({k: v.numpy() for k, v in out.items()} | {'shard0': (beg, end)})
        else:
            assert out is None
            q.put({'shard1': (beg, end)})
    finally:
        dist.destroy_process_group()


def _run(n_steps, ncell, weights):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_steps, ncell, weights, q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        res.update(q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_time_sharding_matches_single_process():
    n_steps, ncell = 11, 7
    res = _run(n_steps, ncell, None)
    full = _field(0, n_steps, ncell)
    for lab in full:
        assert np.array_equal(res[lab], full[lab].numpy())
    assert res['shard0'] == (0, 5) and res['shard1'] == (5, 11)


def test_weighted_and_empty_shards():
    w = np.r_[np.full(2, 50.0), np.full(8, 1.0)]
    res = _run(10, 3, w)
    assert np.array_equal(res['OK'], _field(0, 10, 3)['OK'].numpy())
    assert res['shard0'][1] == res['shard1'][0]
    # one rank may get nothing (fewer steps than ranks)
    res = _run(1, 4, None)
    assert np.array_equal(res['OK'], _field(0, 1, 4)['OK'].numpy())
