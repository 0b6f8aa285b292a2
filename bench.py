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
        mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
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
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.2)

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

    def run_resident(n):
        """n chunks, pipelined one deep: chunk i+1 is prepared and queued while
        chunk i runs; outputs stay in HBM."""
        pend = None
        for _ in range(n):
            nxt = eng.submit_chunk(**kw, **next_chunk(chunks))
            if pend is not None:
                pend.result(to_host=False)
            pend = nxt
        pend.result(to_host=False)

    # pinned host buffers for the end-to-end path (inputs and double-buffered outputs)
    pin_out = [torch.empty((CHUNK_STEPS, NY * NX), dtype=torch.float32).pin_memory()
               for _ in range(2)]
    chunks_e2e = []
    for c in chunks:
        c2 = dict(c)
        c2['data'] = torch.from_numpy(c['data']).pin_memory().numpy()
        chunks_e2e.append(c2)
    copy_stream = torch.cuda.Stream()
    copy_done = [None, None]
    checks = []

    def run_e2e(n):
        """Same pipeline through the public call with HOST inputs; every chunk's
        field is copied to pinned host memory on a copy stream that overlaps the
        next chunk's compute."""
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
            nxt = eng.submit_chunk(**kw, **next_chunk(chunks_e2e))
            if pend is not None:
                drain(pend, k - 1)
            pend = nxt
        drain(pend, n - 1)
        copy_stream.synchronize()
        checks.append(float(pin_out[(n - 1) % 2][0, 0]))

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
    args.warmup = max(args.warmup, 8)
    run_resident(args.warmup)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile_gemm = True
    eng.kernel_events = []
    l0 = eng.total_launches
    ms, ms_wall = timed(run_resident, args.steps)
    launches_timed = eng.total_launches - l0
    eng.collect_profile()
    kernel_events = eng.kernel_events
    eng.profile_gemm = False
    torch.cuda.synchronize()
    engine_stats = dict(eng.stats)

    def summarise(events):
        """Dominant estimate kernel of the timed region: total time, work, rate."""
        by = {}
        for name, bound, work, e0, e1 in events:
            d = by.setdefault(name, dict(bound=bound, work=0.0, ms=0.0, n=0))
            d['work'] += work
            # native submits report the milliseconds between their own CUDA events
            d['ms'] += e0 if e1 is None else e0.elapsed_time(e1)
            d['n'] += 1
        if not by:
            return None
        name = max(by, key=lambda k_: by[k_]['ms'])
        d = by[name]
        d['kernel'] = name
        return d

    dom = summarise(kernel_events)

    run_e2e(min(args.warmup, 2))
    h2d0 = eng.h2d_bytes
    ms_e2e, _ = timed(run_e2e, args.steps)
    h2d_e2e_bytes = eng.h2d_bytes - h2d0
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

    value = world * cell_steps * args.steps / (ms / 1e3)
    e2e_value = world * cell_steps * args.steps / (ms_e2e / 1e3)
    value_dense = world * cell_steps * n_dense / (ms_dense / 1e3)

    if rank == 0:
        peak_f64 = dgemm_peak_tflops(torch)
        hbm_peak, hbm_src = 6650.0, 'fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)'
        mp = ROOT / 'MEASURED_PEAKS.json'
        if mp.exists():
            try:
                hbm_peak = float(json.loads(mp.read_text())['hbm_gbs'])
                hbm_src = 'MEASURED_PEAKS.json hbm_gbs (copy, read + write)'
            except Exception:
                pass
        traffic_tbl = {}
        tf = ROOT / 'profiles' / 'kernel_traffic.json'
        if tf.exists():
            try:
                traffic_tbl = json.loads(tf.read_text())
            except Exception:
                traffic_tbl = {}

        def roofline(d):
            if d is None:
                return None
            avg_ms = d['ms'] / d['n']
            per_launch = d['work'] / d['n']
            if d['bound'] == 'tensor':
                ach = per_launch / (avg_ms / 1e3) / 1e12
                peak, unit = peak_f64, 'TFLOP/s'
                src = ('FP64 tensor: cuBLAS DGEMM 6144^3 via torch.matmul measured live in this '
                       'run (MEASURED_PEAKS.json has no FP64 entry); DMMA issue-rate '
                       'microbenchmark 37.15 TFLOP/s in profiles/microbench')
            else:
                ach = per_launch / (avg_ms / 1e3) / 1e9
                peak, unit, src = hbm_peak, 'GB/s', hbm_src
            t = traffic_tbl.get(d['kernel'], {})
            return {'kernel': 'spx::' + d['kernel'], 'bound': d['bound'], 'achieved': ach,
                    'peak': peak, 'unit': unit, 'frac': ach / peak,
                    'traffic': t.get('dram_bytes_per_launch'),
                    'algorithmic_per_launch': per_launch, 'avg_launch_ms': avg_ms,
                    'launches_timed': d['n'], 'peak_source': src}

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
                'pipeline': 'chunk i+1 is prepared/queued while chunk i runs (engine.submit_chunk)',
                'chunks': '%d distinct chunks of time steps cycled; data-independent per-job '
                          'structures (inverse of the full station system, stations within the '
                          'variogram range of each cell) are cached across chunks' % N_VARIANTS,
                'estimator': ('local (compact-support) estimator: %s' % bool(
                    engine_stats.get('local_rows'))),
                'wall_ms_per_step': ms_wall / args.steps},
            'e2e': {'value': e2e_value, 'unit': 'cell-steps/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(cell_steps * 4),
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches_timed),
            'roofline': roofline(dom),
            'step_breakdown': {'ms_per_step_traced': ms_tr / n_tr, 'entry_points': breakdown},
            'dense_path': {
                'note': 'same workload with the local estimator disabled: every estimate goes '
                        'through the fused variogram-fill + DMMA contraction',
                'value': value_dense, 'unit': 'cell-steps/s', 'ms_per_step': ms_dense / n_dense,
                'roofline': roofline(dense_dom)},
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
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
