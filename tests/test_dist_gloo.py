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


# ---- streamed gather of (time chunk x grid-row chunk) tasks -------------------------------

def test_plan_tasks_covers_the_job_once_and_balances():
    # one row chunk: contiguous blocks of time chunks per rank
    t = sdist.plan_tasks([0, 3, 6, 9, 11], [0, 20], 2)
    assert [x[:4] for x in t] == [(0, 3, 0, 20), (3, 6, 0, 20), (6, 9, 0, 20), (9, 11, 0, 20)]
    assert [x[4] for x in t] == [0, 0, 1, 1]
    # row chunks: the row chunk decides the owner (a rank keeps one geometry)
    t = sdist.plan_tasks([0, 2, 3], [0, 5, 10, 16], 2)
    assert sorted(x[:4] for x in t) == sorted(
        (a, b, c, d) for (a, b) in ((0, 2), (2, 3)) for (c, d) in ((0, 5), (5, 10), (10, 16)))
    assert {x[2]: x[4] for x in t} == {0: 0, 5: 1, 10: 0}
    cover = np.zeros((3, 16), dtype=int)
    for a, b, c, d, _ in t:
        cover[a:b, c:d] += 1
    assert (cover == 1).all()
    # empty chunks are dropped; more ranks than tasks leaves ranks idle
    assert sdist.plan_tasks([0, 0, 4], [0, 8], 4) == [(0, 4, 0, 8, 1)] or \
        len(sdist.plan_tasks([0, 0, 4], [0, 8], 4)) == 1


def _task_field(task, n_cols, lab):
    tb, te, rb, re = task[:4]
    t = torch.arange(tb, te, dtype=torch.float32)[:, None]
    c = torch.arange(rb * n_cols, re * n_cols, dtype=torch.float32)[None, :]
    return (t * 1000 + c) * (1.0 if lab == 'OK' else -0.5)


def _gather_worker(rank, world, port, time_b, row_b, n_cols, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        labels = ['OK', 'IDW_000']
        tasks = sdist.plan_tasks(time_b, row_b, world)
        sg = sdist.StreamedGather(tasks, labels, writer=0, depth=2)
        mine = sg.my_tasks()
        n_t, n_r = time_b[-1], row_b[-1]
        full = {lab: np.full((n_t, n_r * n_cols), np.nan, dtype=np.float32) for lab in labels}
        seen_stats = []

        def put(task, lab, arr):
            tb, te, rb, re = task[:4]
            full[lab].reshape(n_t, n_r, n_cols)[tb:te, rb:re] = arr.reshape(te - tb, re - rb, n_cols)

        def consume(task, lab, buf, st):
            put(task, lab, buf.numpy().copy())
            seen_stats.append(float(st[0, 0]))

        for j in range(sg.n_rounds + 1):
            if 0 <= j - 1 < len(mine):
                task = mine[j - 1]
                slabs = {lab: _task_field(task, n_cols, lab) for lab in labels}
                stats = {lab: torch.full((5, task[1] - task[0]), float(task[0]), dtype=torch.float64)
                         for lab in labels}
                if rank == 0:
                    for lab in labels:
                        put(task, lab, slabs[lab].numpy())
                else:
                    sg.send(slabs, stats)
            if rank == 0 and j >= 1:
                sg.receive_round(j - 1, n_cols, torch.float32, 'cpu', consume)
        sg.flush()
        dist.barrier()
        if rank == 0:
            q.put(dict(full=full, rounds=sg.n_rounds, ring=len(sg._ring or []),
                       received=sg.bytes_received, stats=seen_stats))
        else:
            q.put({})
    finally:
        dist.destroy_process_group()


def _run_gather(time_b, row_b, n_cols, world=2):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, time_b, row_b, n_cols, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        res.update(q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_streamed_gather_time_sharded_matches_single_process():
    """Two gloo ranks, 5 time chunks (2 + 3), two labels: the writer assembles the full
    fields from its own tasks and the slabs received round by round into a 2-slot ring."""
    time_b, row_b, n_cols = [0, 3, 5, 9, 10, 14], [0, 6], 5
    res = _run_gather(time_b, row_b, n_cols)
    for lab in ('OK', 'IDW_000'):
        exp = _task_field((0, 14, 0, 6), n_cols, lab).numpy()
        assert np.array_equal(res['full'][lab], exp)
    assert res['ring'] == 2                      # bounded by the ring, not by the job
    assert res['rounds'] == 3
    assert res["received"] == 2 * 4 * (14 - 5) * 6 * n_cols     # rank 1 owns chunks 2..4


def test_streamed_gather_grid_row_sharded():
    """Few steps, several grid-row chunks: tasks are owned by row chunk."""
    time_b, row_b, n_cols = [0, 2, 3], [0, 3, 7, 9], 4
    res = _run_gather(time_b, row_b, n_cols)
    exp = _task_field((0, 3, 0, 9), n_cols, 'OK').numpy()
    assert np.array_equal(res['full']['OK'], exp)
