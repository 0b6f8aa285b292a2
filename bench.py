#!/usr/bin/env python
"""Benchmark of the gridded-interpolation hot path (BASELINE.json metric:
interpolated cell-steps/s, OK, FP64 arithmetic, f32 store).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path (the compute half of
``SpInterpSteps.interpolate_subset``) over one time chunk of BASELINE config 2:
ordinary kriging of 500 stations with ~20 % missing data (one availability
group per step) onto the full 1000 x 1000 grid, CHUNK_STEPS daily steps per
chunk (the full config is 10 such chunks; the reference's own scheduler splits
the time axis the same way, interp/main.py:652-859).  With N GPUs every rank
processes its own chunk of different time steps (weak scaling, no data-path
collective).

Prints ONE JSON line (see the task contract) with `value` (inputs resident,
outputs left in HBM), `e2e` (host buffers in, pinned host buffers out),
`roofline` of the dominant kernel and `cpu_baseline` (oracle port timed on the
host cores).
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

for _v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
    os.environ.setdefault(_v, '1')          # like the reference (__init__.py:5-10)

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

N_STN = 500
NY = NX = 1000
CHUNK_STEPS = 1250   # one rank's shard of config 2 (10,000 steps) on 8 x B200
MISS = 0.2
VG = '0.1 Nug(0.0) + 0.9 Sph(20000)'
INTERP_ARGS = [('OK', None, 'OK')]
NMRL_PRCN = 2        # decimals the writer rounds to (reference test/test_interp.py:46)
WORKLOAD = ('C2 time shard (1/8 of its 10,000 steps): OK, 500 stations x %d daily steps (20%% '
            'missing, ~1 availability group per step) -> 1000x1000 grid, vg %s'
            % (CHUNK_STEPS, VG))


def make_chunk(rank, variant=0):
    from tests.synth import make_problem
    p = make_problem(2, N_STN, CHUNK_STEPS, NY, NX, cell=1000.0, miss=MISS)
    # same stations / grid on every rank and chunk (one job), different time steps
    rng = np.random.default_rng(1000 + rank + 97 * variant)
    data = rng.gamma(1.0, 5.0, size=(CHUNK_STEPS, N_STN))
    data[rng.random((CHUNK_STEPS, N_STN)) < MISS] = np.nan
    p['data'] = data
    return p


N_VARIANTS = 4   # distinct chunks of time steps cycled through by the timed loops


# ----------------------------------------------------------------- CPU arm

def _cpu_worker(args):
    (data, p, row0, row1, use_ref) = args
    import contextlib
    import io
    if use_ref:
        # the reference's own Cython + Python path, compiled into oracle/_ref
        from oracle import ref_runner
        ref_runner.load()
        case = dict(data=data, stn_xs=p['stn_xs'], stn_ys=p['stn_ys'], cell_xs=p['cell_xs'],
                    cell_ys=p['cell_ys'], grid_shape=p['grid_shape'], interp_args=INTERP_ARGS,
                    vgs=[VG] * data.shape[0], fld_beg_row=row0, fld_end_row=row1)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ref_runner.run_case(case, intrp_dtype=np.float32)
        return data.shape[0] * (row1 - row0) * p['grid_shape'][1], time.perf_counter() - t0
    from oracle import spinterp_oracle as orc
    t0 = time.perf_counter()
    flds, _ = orc.interp_chunk(
        data, p['stn_xs'], p['stn_ys'], p['cell_xs'], p['cell_ys'], p['grid_shape'],
        INTERP_ARGS, vgs=[VG] * data.shape[0], fld_beg_row=row0, fld_end_row=row1,
        intrp_dtype=np.float32, faithful=True)
    return data.shape[0] * (row1 - row0) * p['grid_shape'][1], time.perf_counter() - t0


def cpu_kind():
    """'reference' when oracle/_ref (the reference compiled by oracle/build_ref.py) is
    present, else 'port' (the NumPy oracle with the reference's loop structure)."""
    try:
        from oracle import ref_runner
        return 'reference' if ref_runner.available() else 'port'
    except Exception:
        return 'port'


def cpu_sample(p, n_cores, steps_per_core=1, rows=8, kind=None):
    """The reference's CPU path, multiprocess over time chunks like
    interp/main.py:141-153 (one fresh SpInterpSteps per task); returns
    (cell_steps, seconds, n_steps, rows)."""
    import multiprocessing as mp
    use_ref = (kind or cpu_kind()) == 'reference'
    n_steps = n_cores * steps_per_core
    tasks = [(p['data'][i * steps_per_core:(i + 1) * steps_per_core], p, 0, rows, use_ref)
             for i in range(n_cores)]
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    with ctx.Pool(n_cores) as pool:
        res = pool.map(_cpu_worker, tasks, chunksize=1)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall, n_steps, rows


CPU_DESC = {'reference': "the reference's own interp/steps.py + cyth/interpmthds.pyx compiled "
                         "into oracle/_ref, Pool(%d) over time chunks",
            'port': 'oracle port with the reference loop structure, Pool(%d) over time chunks'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_cores = len(os.sched_getaffinity(0))
    kind = cpu_kind()
    p = make_chunk(0)
    vals = []
    for it in range(args.warmup + args.steps):
        cs, wall, n_steps, rows = cpu_sample(p, n_cores, steps_per_core=1, rows=64, kind=kind)
        if it >= args.warmup:
            vals.append((cs, wall))
    cs = sum(v[0] for v in vals)
    wall = sum(v[1] for v in vals)
    value = cs / wall
    sample = ('%d steps x %d grid rows x %d cols per bench step (same stations, missingness and '
              'variogram as the GPU arm); %s' % (n_steps, rows, NX, CPU_DESC[kind] % n_cores))
    line = {
        'impl': 'reference', 'metric': 'interpolated cell-steps/s (OK, FP64, f32 store)',
        'value': value, 'unit': 'cell-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'cell-steps/s', 'cores': n_cores, 'kind': kind,
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------- GPU arm

class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region.  Sampled in-process through
    NVML (pynvml): forking nvidia-smi from a process that holds a CUDA context and ~10 GB
    of pinned memory stalls the main thread for milliseconds -- longer than a bench step.
    nvidia-smi is only the fallback when pynvml is missing."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.nvml = None
        self._max_sm = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = index
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                ids = [v.strip() for v in vis.split(',') if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    idx = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        nv = self.nvml
        sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
        if self._max_sm is None:                       # constant: asked once
            self._max_sm = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        mx = self._max_sm
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)]
        return [str(sm), str(mx)] + ['Active' if (r & b) else 'Not Active' for b in bits]

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(
                        ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                         '--format=csv,noheader,nounits'], capture_output=True, text=True,
                        timeout=5)
                    parts = [x.strip() for x in out.stdout.strip().split(',')]
                    if len(parts) >= 6:
                        self.samples.append(parts)
            except Exception:
                pass
            # 10 samples per second: every NVML query takes the driver's lock, which on an
            # 8-GPU box holds up the kernel launches of all ranks for a moment; the timed
            # regions of a run add up to seconds, so the median clock is still well sampled
            self.stop_flag.wait(0.1 if self.nvml is not None else 0.25)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names)
                   if any(s[2 + k].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(self.samples),
                'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def dgemm_peak_tflops(torch, n=6144, reps=4):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = float('inf')
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / best / 1e9


def hbm_peak():
    hbm, src = 6650.0, 'fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)'
    mp = ROOT / 'MEASURED_PEAKS.json'
    if mp.exists():
        try:
            hbm = float(json.loads(mp.read_text())['hbm_gbs'])
            src = 'MEASURED_PEAKS.json hbm_gbs (copy, read + write)'
        except Exception:
            pass
    return hbm, src


def summarise_events(events):
    """Per kernel: total time, work, launches; returns (dominant kernel dict, all)."""
    by = {}
    for name, bound, work, e0, e1 in events:
        d = by.setdefault(name, dict(bound=bound, work=0.0, ms=0.0, n=0))
        d['work'] += work
        # native submits report the milliseconds between their own CUDA events
        d['ms'] += e0 if e1 is None else e0.elapsed_time(e1)
        d['n'] += 1
    if not by:
        return None, {}
    name = max(by, key=lambda k_: by[k_]['ms'])
    by[name]['kernel'] = name
    return by[name], by


def roofline_of(d, peak_f64, traffic_tbl=None):
    if d is None:
        return None
    hbm, hbm_src = hbm_peak()
    avg_ms = d['ms'] / d['n']
    per_launch = d['work'] / d['n']
    if d['bound'] == 'tensor':
        ach = per_launch / (avg_ms / 1e3) / 1e12
        peak, unit = peak_f64, 'TFLOP/s'
        src = ('FP64 tensor: cuBLAS DGEMM 6144^3 via torch.matmul measured live in this '
               'run (MEASURED_PEAKS.json has no FP64 entry); DMMA issue-rate '
               'microbenchmark 37.15 TFLOP/s in profiles/microbench')
    elif d['bound'] == 'alu':
        ach = per_launch / (avg_ms / 1e3) / 1e12
        peak, unit = 33.9, 'TFLOP/s'
        src = 'FP64 ALU: DFMA issue-rate microbenchmark 33.9 TFLOP/s (profiles/microbench)'
    else:
        ach = per_launch / (avg_ms / 1e3) / 1e9
        peak, unit, src = hbm, 'GB/s', hbm_src
    t = (traffic_tbl or {}).get(d['kernel'], {})
    return {'kernel': 'spx::' + d['kernel'], 'bound': d['bound'], 'achieved': ach,
            'peak': peak, 'unit': unit, 'frac': ach / peak,
            'traffic': t.get('dram_bytes_per_launch'),
            'traffic_source': ('profiles/kernel_traffic.json: one `ncu --set full` capture of '
                               'this kernel on this workload (static, not re-measured per run)'
                               if t else None),
            'algorithmic_per_launch': per_launch, 'avg_launch_ms': avg_ms,
            'launches_timed': d['n'], 'peak_source': src}


CONFIG_NOTE = {
    'C1': 'C1: test_interp-style run, OK + IDW(2), 100 stations x 365 daily steps -> 200x200 grid',
    'C3': 'C3 time chunk: EDK with an elevation drift, 300 stations x 250 of its 5,000 steps -> '
          '2000x2000 grid, one variogram string per step',
    'C4': 'C4 time chunk: IDW exponents {1,2,3,5}, 2,000 stations x 1,000 of its 20,000 hourly '
          'steps -> 1000x1000 grid',
    'C5': 'C5 time chunk: SK + OK, 1,000 stations x 100 of its 2,000 steps -> 4000x4000 grid with '
          'an elliptic cell mask (~49 % of the cells)',
}


def run_config(args):
    """`--config C1|C3|C4|C5`: one bench step = one time chunk of that BASELINE.json
    configuration on every rank (weak scaling, as for the default C2 line)."""
    import torch
    import torch.distributed as dist
    from spinterps_b200.engine import ChunkEngine
    from tests.synth import CONFIG_CHUNK, config_problem

    cfg = args.config
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    chunk = CONFIG_CHUNK[cfg]
    variants = []
    for v in range(2):
        p, kw = config_problem(cfg, chunk, seed_shift=1 + 16 * rank + v)
        variants.append((p, kw))
    base = {k: v for k, v in variants[0][0].items() if k != 'data'}
    n_labels = len(variants[0][1]['interp_args'])
    G = int(variants[0][0]['cell_xs'].size)
    cell_steps = n_labels * G * chunk
    eng = ChunkEngine()
    counter = [0]

    def submit(**extra):
        counter[0] += 1
        p, kw = variants[counter[0] % 2]
        return eng.submit_chunk(p['data'], intrp_dtype=np.float32, **base, **kw, **extra)

    def run_resident(n):
        pend = None
        for _ in range(n):
            nxt = submit()
            if pend is not None:
                pend.result(to_host=False)
            pend = nxt
        pend.result(to_host=False)

    def run_e2e(n):
        for _ in range(n):
            submit(round_decimals=NMRL_PRCN, field_stats=True).result(to_host=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    run_resident(max(args.warmup, 3))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile_gemm = True
    eng.kernel_events = []
    l0 = eng.total_launches
    ms = timed(run_resident, args.steps)
    launches = eng.total_launches - l0
    eng.collect_profile()
    dom, by = summarise_events(eng.kernel_events)
    eng.profile_gemm = False
    stats = dict(eng.stats)
    n_e2e = max(1, min(args.steps, 3))
    run_e2e(1)
    h2d0 = eng.h2d_bytes
    d2h0 = eng._dl.d2h_bytes if eng._dl is not None else 0
    ms_e2e = timed(run_e2e, n_e2e)
    d2h = (eng._dl.d2h_bytes - d2h0) if eng._dl is not None else n_e2e * cell_steps * 4
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
        peak_f64 = dgemm_peak_tflops(torch)
        line = {
            'metric': 'interpolated cell-steps/s (FP64 arithmetic, f32 store)',
            'value': world * cell_steps * args.steps / (ms / 1e3), 'unit': 'cell-steps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': CONFIG_NOTE[cfg], 'config': cfg, 'chunk_steps': chunk,
                       'cells': G, 'labels': [a[2] for a in variants[0][1]['interp_args']],
                       'parallelism': 'time-sharded x%d' % world,
                       'l2': 'each step writes %.1f GB of fields (>> 126 MB L2)'
                             % (cell_steps * 4 / 1e9)},
            'e2e': {'value': world * cell_steps * n_e2e / (ms_e2e / 1e3), 'unit': 'cell-steps/s',
                    'h2d_bytes_per_step': int((eng.h2d_bytes - h2d0) / n_e2e),
                    'd2h_bytes_per_step': int(d2h / n_e2e), 'ms_per_step': ms_e2e / n_e2e,
                    'note': 'submit_chunk(round_decimals=2, field_stats=True).result(): '
                            'synchronous, the rounded fields come back as host arrays'},
            'gpu_launches': int(launches),
            'roofline': roofline_of(dom, peak_f64),
            'kernels': {k_: {'ms_per_step': v['ms'] / args.steps, 'launches_per_step': v['n'] / args.steps,
                             'bound': v['bound']} for k_, v in by.items()},
            'engine_stats': {k_: (int(v_) if isinstance(v_, (int, np.integer)) else None)
                             for k_, v_ in stats.items() if isinstance(v_, (int, np.integer))},
            'cpu_baseline': None, 'clocks': sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from spinterps_b200.engine import ChunkEngine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    p = make_chunk(rank)
    cpu_res = None
    if world == 1:
        # CPU baseline first: fork the worker pool before CUDA is initialised
        n_cores = len(os.sched_getaffinity(0))
        cpu_res = cpu_sample(p, n_cores, steps_per_core=1, rows=64)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = ChunkEngine()
    vgs = [VG] * CHUNK_STEPS
    cell_steps = CHUNK_STEPS * NY * NX

    kw = dict(interp_args=INTERP_ARGS, vgs=vgs, intrp_dtype=np.float32)

    # consecutive bench steps process DIFFERENT time steps (data, missingness and
    # availability groups differ); stations, grid and variogram are the job's
    chunks = [p] + [make_chunk(rank, v) for v in range(1, N_VARIANTS)]
    counter = [0]

    def next_chunk(pool):
        counter[0] += 1
        return pool[counter[0] % len(pool)]

    PIPE_DEPTH = 2

    def run_resident(n):
        """n chunks, pipelined: up to PIPE_DEPTH chunks are prepared and queued ahead of the
        one whose result is taken (the host runs ahead of the GPU, so a slow host moment --
        eight ranks share the host's cores -- does not starve the device); outputs stay in
        HBM."""
        pend = collections.deque()
        for _ in range(n):
            if len(pend) == PIPE_DEPTH:
                pend.popleft().result(to_host=False)
            pend.append(eng.submit_chunk(**kw, **next_chunk(chunks)))
        while pend:
            pend.popleft().result(to_host=False)

    # ---- end-to-end path: host inputs (pinned), results back in host memory ------------
    # The chunk goes through the public call with the output stage of the reference's writer
    # (fields rounded to NMRL_PRCN decimals + per-step statistics on the device, like
    # SpInterpMain.interpolate()).  Default transport: the rounded f32 field crosses PCIe as
    # 16-bit codes and is decoded to the identical floats by host threads
    # (spinterps_b200/transfer.py); `raw`: the f32 field itself into pinned memory (round 1).
    from spinterps_b200.transfer import PackedDownloader
    import concurrent.futures
    chunks_e2e = []
    for c in chunks:
        c2 = dict(c)
        c2['data'] = torch.from_numpy(c['data']).pin_memory().numpy()
        chunks_e2e.append(c2)
    copy_stream = torch.cuda.Stream()
    checks = []
    n_dec = int(os.environ.get('SPX_DECODE_THREADS', str(max(1, min(16, len(os.sched_getaffinity(0)) // world)))))
    host_out = [np.empty((CHUNK_STEPS, NY * NX), dtype=np.float32) for _ in range(2)]
    decoder = concurrent.futures.ThreadPoolExecutor(
        max_workers=1, initializer=lambda: torch.cuda.set_device(local_rank))
    kw_e2e = dict(kw, round_decimals=NMRL_PRCN, field_stats=True)

    def run_e2e(n, decode=False):
        """Pipeline: submit chunk i+1 | output stage + encode + D2H of chunk i, through the
        engine's public PendingChunk.start_packed / finish_packed.  At the end of a step its
        result sits in (pinned) host memory in the lossless compact form the writer consumes
        (transfer.DeltaField / PackedField) together with the per-step statistics;
        decode=True additionally rebuilds the whole f32 field in host memory (host threads,
        one chunk behind)."""
        futs = [None, None]

        def land(pend, handle, k):
            out, _ = pend.finish_packed(handle)
            pf = out['OK']
            if decode:
                pf.decode(host_out[k % 2], n_dec)
            checks.append(int(pf.nbytes))
            assert pend.field_stats is not None
            pf.release()

        def drain(pend, k):
            if futs[k % 2] is not None:
                futs[k % 2].result()                 # slot (and host buffer) k % 2 free again
            futs[k % 2] = decoder.submit(land, pend, pend.start_packed(), k)
        pend = None
        for k in range(n):
            nxt = eng.submit_chunk(**kw_e2e, **next_chunk(chunks_e2e))
            if pend is not None:
                drain(pend, k - 1)
            pend = nxt
        drain(pend, n - 1)
        for f in futs:
            if f is not None:
                f.result()

    pin_out = []
    copy_done = [None, None]

    def run_e2e_raw(n):
        """Round-1 transport for comparison: the rounded f32 field copied to pinned host
        memory on a copy stream that overlaps the next chunk's compute."""
        if not pin_out:
            pin_out.extend(torch.empty((CHUNK_STEPS, NY * NX), dtype=torch.float32).pin_memory()
                           for _ in range(2))

        def drain(pend, k):
            flds, _ = pend.result(to_host=False)
            buf = pin_out[k % 2]
            if copy_done[k % 2] is not None:
                copy_done[k % 2].synchronize()
            copy_stream.wait_event(pend.done_event)
            with torch.cuda.stream(copy_stream):
                buf.copy_(flds['OK'], non_blocking=True)
                flds['OK'].record_stream(copy_stream)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            copy_done[k % 2] = ev
        pend = None
        for k in range(n):
            nxt = eng.submit_chunk(**kw_e2e, **next_chunk(chunks_e2e))
            if pend is not None:
                drain(pend, k - 1)
            pend = nxt
        drain(pend, n - 1)
        copy_stream.synchronize()
        checks.append(float(pin_out[(n - 1) % 2][0, 0]))

    def d2h_ceiling_gbs(nbytes=1 << 30):
        """Bare cudaMemcpyAsync device -> pinned host of 1 GiB, best of 3."""
        src = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
        dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        best = float('inf')
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del src, dst
        return nbytes / best / 1e6

    def timed(fn, n):
        import gc
        gc.collect()
        gc.disable()        # a generation-2 collection inside a 30 ms region is a 10 % outlier
        try:
            return _timed(fn, n)
        finally:
            gc.enable()

    def _timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn(n)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms_dev = e0.elapsed_time(e1)
        ms_wall = 1e3 * (time.perf_counter() - t0)
        ms = max(ms_dev, 0.0)
        if world > 1:
            t = torch.tensor([ms, ms_wall], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_wall = float(t[0]), float(t[1])
        return ms, ms_wall

    # Steady state needs more than 3 chunks: the first chunk of a job builds the inverse of
    # the full station system and the near-station tables (cached afterwards), the upload
    # arena ring has 4 slots that are allocated on first use, the caching allocator has
    # to see the 5 GB field blocks once.  Untimed warm-up is therefore at least 8 chunks.
    n_prime = max(args.warmup, 8)
    if world > 1:
        # finish the (lazy) NCCL communicator set-up before anything is timed
        t = torch.ones(1, device='cuda')
        dist.all_reduce(t)
        dist.barrier()
        torch.cuda.synchronize()
        n_prime = max(n_prime, 12)
    run_resident(n_prime)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile_gemm = True
    eng.kernel_events = []
    eng.solve_ms = []
    l0 = eng.total_launches
    ms, ms_wall = timed(run_resident, args.steps)
    launches_timed = eng.total_launches - l0
    eng.collect_profile()
    kernel_events = eng.kernel_events
    solve_ms = list(eng.solve_ms)
    sparse_job = next((j for j in eng._fast_jobs.values() if j['cfg'].sparse.n_comp > 0), None)
    eng.profile_gemm = False
    torch.cuda.synchronize()
    engine_stats = dict(eng.stats)

    def summarise(events):
        return summarise_events(events)[0]

    dom = summarise(kernel_events)

    # the pipelined loop keeps up to three 5 GB fields alive and a few more pinned blocks than
    # the resident loop: let both caching allocators see that before anything is timed (with
    # several ranks on one host the first cudaMalloc / cudaHostAlloc calls are slow)
    n_prime_e2e = max(args.warmup, 8)
    run_e2e(n_prime_e2e)
    dl = eng._dl
    h2d0 = eng.h2d_bytes
    d2h0 = dl.d2h_bytes
    ms_e2e, _ = timed(run_e2e, args.steps)
    h2d_e2e_bytes = eng.h2d_bytes - h2d0
    d2h_e2e_bytes = dl.d2h_bytes - d2h0
    e2e_codec, e2e_fallbacks, u16_stride = dl.codec, dl.fallbacks, dl.stride
    # ... with the whole f32 field rebuilt in host memory inside the timed region
    n_dcd = max(2, min(args.steps, 5))
    run_e2e(2, decode=True)
    ms_e2e_dec, _ = timed(lambda n: run_e2e(n, decode=True), n_dcd)
    # the same pipeline with the 16-bit codes of the earlier transport
    n_u16 = max(2, min(args.steps, 4))
    ms_e2e_u16 = None
    if e2e_codec != 'u16':
        eng.transport = 'u16'
        run_e2e(3)
        ms_e2e_u16, _ = timed(run_e2e, n_u16)
        eng.transport = None
        eng._dl = dl = None
    # the same pipeline with the round-1 transport (f32 field into pinned memory)
    n_raw = max(2, min(args.steps, 4))
    run_e2e_raw(2)
    ms_e2e_raw, _ = timed(run_e2e_raw, n_raw)
    pin_out.clear()
    d2h_peak = d2h_ceiling_gbs() if rank == 0 else None
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # per-entry-point breakdown of one resident step (events around every library call;
    # separate from the timed region above)
    eng.trace = []
    eng.trace_launches = True
    n_tr = max(2, min(args.steps, 4))
    ms_tr, _ = timed(run_resident, n_tr)
    eng.trace_launches = False
    breakdown = {k_: {'ms_per_step': round(v['ms'] / n_tr, 4), 'calls_per_step': v['n'] / n_tr}
                 for k_, v in sorted(eng.trace_summary().items(), key=lambda kv: -kv[1]['ms'])}
    eng.trace = []

    # the tensor-core contraction on the same workload (local estimator off), so that
    # both estimators are on record
    eng.local_support = False
    run_resident(1)
    eng.profile_gemm = True
    eng.kernel_events = []
    n_dense = max(2, min(args.steps, 3))
    ms_dense, _ = timed(run_resident, n_dense)
    eng.collect_profile()
    dense_dom = summarise(eng.kernel_events)
    eng.profile_gemm = False
    eng.local_support = True

    # ---- the one collective of the design: finished slabs -> writer rank (N > 1) --------
    gather = None
    if world > 1:
        fld = eng.submit_chunk(**kw_e2e, **next_chunk(chunks_e2e)).result(to_host=False)[0]['OK']
        nbytes = fld.numel() * 4
        ring = [torch.empty_like(fld) for _ in range(2)] if rank == 0 else None

        def gather_once(_n):
            # the protocol of SpInterpMain.interpolate (dist.StreamedGather): the writer
            # receives one slab at a time into a 2-slot ring
            if rank == 0:
                for r in range(1, world):
                    dist.recv(ring[r % 2], src=r)
            else:
                dist.send(fld, dst=0)
        gather_once(1)
        ms_g, _ = timed(gather_once, 1)
        gather = {'ms': ms_g, 'bytes_into_writer': int((world - 1) * nbytes),
                  'gbs_into_writer': (world - 1) * nbytes / ms_g / 1e6,
                  'nvlink_nominal_gbs_per_direction': 900.0,
                  'note': 'NCCL send/recv of every other rank\'s rounded f32 chunk (5 GB each) to the '
                          'writer GPU, one slab at a time into a 2-slot ring (no download)'}
        del fld, ring

    value = world * cell_steps * args.steps / (ms / 1e3)
    e2e_value = world * cell_steps * args.steps / (ms_e2e / 1e3)
    value_dense = world * cell_steps * n_dense / (ms_dense / 1e3)

    if rank == 0:
        peak_f64 = dgemm_peak_tflops(torch)
        traffic_tbl = {}
        tf = ROOT / 'profiles' / 'kernel_traffic.json'
        if tf.exists():
            try:
                traffic_tbl = json.loads(tf.read_text())
            except Exception:
                traffic_tbl = {}

        def roofline(d):
            return roofline_of(d, peak_f64, traffic_tbl)

        cpu_baseline = None
        if cpu_res is not None:
            cs, wall, n_steps_s, rows_s = cpu_res
            kind = cpu_kind()
            sample = ('%d steps x %d grid rows x %d cols (same stations, missingness, variogram); %s'
                      % (n_steps_s, rows_s, NX, CPU_DESC[kind] % n_cores))
            cpu_baseline = {'value': cs / wall, 'unit': 'cell-steps/s', 'cores': n_cores,
                            'kind': kind, 'sample': sample}
        h2d = int(round(h2d_e2e_bytes / max(args.steps, 1)))   # counted by the engine
        line = {
            'metric': 'interpolated cell-steps/s (OK, FP64, f32 store)',
            'value': value, 'unit': 'cell-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': WORKLOAD, 'n_stations': N_STN, 'chunk_steps': CHUNK_STEPS,
                'grid': [NY, NX], 'missing': MISS, 'parallelism': 'time-sharded x%d' % world,
                'l2': 'each step writes a %.1f GB field (>> 126 MB L2) between reuses'
                      % (cell_steps * 4 / 1e9),
                'pipeline': 'up to %d chunks are prepared / queued ahead of the one whose result is '
                            'taken (engine.submit_chunk)' % PIPE_DEPTH,
                'chunks': '%d distinct chunks of time steps cycled; data-independent per-job '
                          'structures (inverse of the full station system, stations within the '
                          'variogram range of each cell) are cached across chunks' % N_VARIANTS,
                'estimator': ('local (compact-support) estimator: %s' % bool(
                    engine_stats.get('local_rows'))),
                'wall_ms_per_step': ms_wall / args.steps,
                'warmup_chunks_run': n_prime, 'e2e_warmup_chunks_run': n_prime_e2e},
            'e2e': {'value': e2e_value, 'unit': 'cell-steps/s',
                    'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': int(round(d2h_e2e_bytes / max(args.steps, 1))),
                    'ms_per_step': ms_e2e / args.steps,
                    'transport': 'output stage of the writer on the device (np.round to %d '
                                 'decimals, per-step statistics); the rounded f32 field crosses '
                                 'PCIe in a lossless compact form (codec %r: first / second '
                                 'differences of np.round\'s integer lattice along each row, '
                                 'bit-packed per 8 cells, produced by the same single pass over '
                                 'the field; size depends on the field, see '
                                 'd2h_bytes_per_cell_step) and lands in pinned host '
                                 'memory as transfer.DeltaField, which the writer decodes one '
                                 'step at a time' % (NMRL_PRCN, e2e_codec),
                    'd2h_bytes_per_cell_step': d2h_e2e_bytes / max(args.steps, 1) / cell_steps,
                    'fields_sent_through_fallback_codec': int(e2e_fallbacks),
                    'u16_transport': None if ms_e2e_u16 is None else {
                        'value': world * cell_steps * n_u16 / (ms_e2e_u16 / 1e3),
                        'ms_per_step': ms_e2e_u16 / n_u16,
                        'd2h_bytes_per_step': int(CHUNK_STEPS * (16 + 2 * u16_stride)),
                        'note': 'same pipeline, 16-bit codes per row (2 bytes per cell-step)'},
                    'decoded_f32': {
                        'value': world * cell_steps * n_dcd / (ms_e2e_dec / 1e3),
                        'ms_per_step': ms_e2e_dec / n_dcd, 'decode_threads': n_dec,
                        'note': 'same pipeline plus the decode of the WHOLE field to f32 in host '
                                'memory inside the timed region'},
                    'd2h_gbs': d2h_e2e_bytes / max(ms_e2e, 1e-9) / 1e6,
                    'd2h_ceiling_gbs': d2h_peak,
                    'raw_f32_transport': {
                        'value': world * cell_steps * n_raw / (ms_e2e_raw / 1e3),
                        'ms_per_step': ms_e2e_raw / n_raw,
                        'd2h_bytes_per_step': int(cell_steps * 4),
                        'note': 'same pipeline, f32 field copied to pinned host memory'}},
            'solve_phase': {
                'ms_per_step': (sum(solve_ms) / len(solve_ms)) if solve_ms else None,
                'method': ('sparse covariance form: the variogram is constant beyond its range, '
                           'the OK matrix is F 11\' - C with C block diagonal over %d clusters of '
                           'stations closer than the range (largest: %d stations); one warp per '
                           'time step, O(n_stn) (spx_krige_sparse_ok_dev)'
                           % (sparse_job['cfg'].sparse.n_comp, sparse_job['cfg'].sparse.max_size))
                if sparse_job is not None else
                'Ut = Bt.G on the FP64 tensor cores + LDL^T downdate of one r x r block per '
                'availability group (spx_krige_downdate_dev)',
                'note': 'CUDA events around the solve phase of every timed chunk (uploads -> '
                        'coefficients ready); SPX_SPARSE_SOLVE=0 selects the downdate'},
            'gpu_launches': int(launches_timed),
            'roofline': roofline(dom),
            'step_breakdown': {'ms_per_step_traced': ms_tr / n_tr, 'entry_points': breakdown,
                               'note': 'separate traced pass through the Python-planned path '
                                       '(downdated dense systems), events around every entry '
                                       'point; the timed region above uses the native submit'},
            'dense_path': {
                'note': 'same workload with the local estimator disabled: every estimate goes '
                        'through the fused variogram-fill + DMMA contraction',
                'value': value_dense, 'unit': 'cell-steps/s', 'ms_per_step': ms_dense / n_dense,
                'roofline': roofline(dense_dom)},
            'gather': gather,
            'cpu_baseline': cpu_baseline,
            'clocks': sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C2', choices=['C1', 'C2', 'C3', 'C4', 'C5'],
                    help='BASELINE.json configuration (default C2: the headline workload)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.config != 'C2':
        run_config(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
