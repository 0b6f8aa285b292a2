"""GPU parity of the chunk engine (C-ABI group 3 driven by engine.py) against
the golden fixtures of the reference and against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): kriging 1e-9 relative, IDW 1e-12
relative, nearest neighbour exact; NaN patterns identical."""
import numpy as np
import pytest

from oracle import spinterp_oracle as orc
from tests.golden_util import load_case, rel_err
from tests.synth import VG_C1, make_problem

pytestmark = pytest.mark.gpu

KRG_TOL = 1e-9
IDW_TOL = 1e-12


@pytest.fixture(scope='module', params=['auto', 'dense'])
def eng(request):
    """'auto': the engine picks its estimator (local compact-support estimator for
    Nug + Sph / Lin variograms with few stations in range, tensor-core contraction
    otherwise); 'dense': always the contraction."""
    from spinterps_b200.engine import ChunkEngine
    e = ChunkEngine()
    if request.param == 'dense':
        e.local_support = False
    return e


def _tol(label):
    if label.startswith('IDW'):
        return IDW_TOL
    if label.startswith('NNB'):
        return 0.0
    return KRG_TOL


def _floor(ref):
    """Relative errors are taken against max(|ref|, 1 % of the field's largest
    value): estimates are weighted sums of data of that magnitude, so cells whose
    value cancels to ~0 carry the same ABSOLUTE rounding error as the others."""
    return max(1e-3, 0.01 * float(np.nanmax(np.abs(ref)))) if np.isfinite(ref).any() else 1e-3


PARITY_REPORT = []      # written to gpurun_out/parity_report.json at the end of the session


def _report(name, lab, got, ref, floor):
    """Besides the floored relative error the tests assert on: the UN-floored maximum
    relative error (over cells with |ref| > 0) and the share of cells whose |ref| lies
    below the floor, i.e. for which the floor matters at all."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    m = np.isfinite(ref) & (ref != 0)
    unfl = float((np.abs(got[m] - ref[m]) / np.abs(ref[m])).max()) if m.any() else 0.0
    fin = np.isfinite(ref)
    PARITY_REPORT.append(dict(
        case=name, label=lab, floor=float(floor), rel_err_floored=rel_err(got, ref, floor),
        rel_err_unfloored=unfl,
        cells_below_floor=float((np.abs(ref[fin]) < floor).mean()) if fin.any() else 0.0,
        max_abs_err=float(np.abs(got[fin] - ref[fin]).max()) if fin.any() else 0.0,
        field_max=float(np.abs(ref[fin]).max()) if fin.any() else 0.0))


def _check(got, exp, name):
    for lab, ref in exp.items():
        e = rel_err(got[lab], ref, _floor(ref))
        _report(name, lab, got[lab], ref, _floor(ref))
        assert e <= _tol(lab), (name, lab, e)


@pytest.mark.parametrize('name', ['a_ok_idw_nnb', 'c_edk_drift', 'd_sk_ok_mask_rows', 'e_nrst',
                                  'g_idw_only', 'h_pie'])
def test_golden_cases(eng, name):
    case, outs = load_case(name)
    got, _ = eng.interp_chunk(intrp_dtype=np.float64, **case)
    assert set(got) == set(outs)
    _check(got, outs, name)


def test_golden_groups_flags(eng):
    """Case b: several availability groups, three variograms (one nugget-only),
    low-value steps, a single-station step, a step without stations, cut-offs and
    the OK estimation variance."""
    case, outs = load_case('b_ok_groups_flags')
    got, prob = eng.interp_chunk(intrp_dtype=np.float64, **case)
    assert set(got) == set(outs)
    _check(got, outs, 'b')
    assert prob == [10]          # the step without any station


def test_golden_vg_families(eng):
    """Every variogram family; the Gau / Pow systems are ill-conditioned
    (cond ~1e8+), LU and the reference's SVD pinv then agree to ~cond*eps only."""
    case, outs = load_case('f_vg_families')
    got, _ = eng.interp_chunk(intrp_dtype=np.float64, **case)
    ref = outs['OK']
    errs = [rel_err(got['OK'][t], ref[t]) for t in range(ref.shape[0])]
    print('per-step rel err', errs)
    for t in (0, 1, 2, 4, 5, 6):  # cond(A) <= 6e3 for these systems
        assert errs[t] <= KRG_TOL, (t, errs[t])
    # step 3 (Hol): cond(A) = 7.2e5 and strongly cancelling weights -- NumPy's own
    # pinv evaluated as gemm instead of the reference's per-cell gemv already
    # differs from the golden field by 8.7e-10, LAPACK solve by 3.7e-9
    assert errs[3] <= 2e-8, errs[3]


def test_float32_store(eng):
    case, outs = load_case('a_ok_idw_nnb')
    got, _ = eng.interp_chunk(intrp_dtype=np.float32, **case)
    for lab, ref in outs.items():
        assert got[lab].dtype == np.float32
        assert np.array_equal(got[lab], ref.astype(np.float32)) or \
            rel_err(got[lab], ref.astype(np.float32)) <= 2e-7


@pytest.mark.parametrize('n_stn,n_steps,ny,nx,miss', [
    (120, 40, 64, 64, 0.2),      # many availability groups
    (100, 37, 50, 41, 0.0),      # one group, ragged tile sizes
    (259, 300, 30, 33, 0.05),    # more than one M-tile of rows, K not a multiple of 8
    (600, 30, 20, 20, 0.21),     # ~126 +- 10 missing stations per step: every tile shape
                                 # of the register-resident downdate kernel
])
def test_seeded_vs_oracle(eng, n_stn, n_steps, ny, nx, miss):
    p = make_problem(11, n_stn, n_steps, ny, nx, cell=1000.0 * 200 / max(ny, nx), miss=miss)
    args = [('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0), ('IDW', None, 'IDW_001', 3.0),
            ('NNB', None, 'NNB')]
    kw = dict(interp_args=args, vgs=[VG_C1] * n_steps, **p)
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    got, _ = eng.interp_chunk(intrp_dtype=np.float64, **kw)
    _check(got, exp, 'seeded')


@pytest.mark.parametrize('knob', ['SPX_DD_PIVOT=1', 'SPX_DD_SMEM=1', 'SPX_DD_LU=0',
                                  'SPX_DD_LU_FAIL=1'])
def test_downdate_kernel_variants(knob):
    """The downdated solves run a blocked LU without pivoting (definite S) by default;
    behind it: Gauss-Jordan in registers without / with pivoting (the latter also as the
    repair pass for systems the LU flags -- SPX_DD_LU_FAIL makes it flag every other
    system) and LU in shared memory with pivoting.  The environment knobs force each of
    them; same parity bar."""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    k, v = knob.split('=')
    env[k] = v
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_engine.py', '-q', '-x',
                        '-m', 'gpu', '-k', 'test_seeded_vs_oracle and auto'],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '4 passed' in r.stdout, r.stdout[-500:]


def test_linearity_and_constant_field(eng):
    """Size-independent properties: kriging / IDW weights sum to one, so a
    constant data field is reproduced; estimates are linear in the data."""
    p = make_problem(5, 80, 6, 40, 40, cell=2500.0)
    args = [('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)]
    vg = [VG_C1] * 6
    base = dict(p)
    d1 = base.pop('data')
    d2 = np.random.default_rng(1).gamma(1.0, 5.0, size=d1.shape)
    r1, _ = eng.interp_chunk(d1, interp_args=args, vgs=vg, intrp_dtype=np.float64, **base)
    r2, _ = eng.interp_chunk(d2, interp_args=args, vgs=vg, intrp_dtype=np.float64, **base)
    r3, _ = eng.interp_chunk(2.0 * d1 + 0.5 * d2, interp_args=args, vgs=vg,
                             intrp_dtype=np.float64, **base)
    for lab in ('OK', 'IDW_000'):
        np.testing.assert_allclose(r3[lab], 2.0 * r1[lab] + 0.5 * r2[lab], rtol=1e-10, atol=1e-10)
    rc, _ = eng.interp_chunk(np.full_like(d1, 7.25), interp_args=args, vgs=vg,
                             intrp_dtype=np.float64, **base)
    for lab in ('OK', 'IDW_000'):
        np.testing.assert_allclose(rc[lab], 7.25, rtol=1e-11)


def test_nnb_candidate_lists_and_full_scan_fallback(eng):
    """Nearest neighbour through candidate lists (16 nearest stations per cell):
    with 85 % missing data most cells have no available candidate and take the
    full-scan branch; indices must equal np.argmin of the oracle exactly."""
    p = make_problem(21, 70, 12, 37, 29, cell=3000.0, miss=0.85)
    p['data'][3, :] = np.nan
    p['data'][3, 5] = 1.5                      # a single-station step
    args = [('NNB', None, 'NNB')]
    exp, _ = orc.interp_chunk(interp_args=args, intrp_dtype=np.float64, **p)
    got, _ = eng.interp_chunk(interp_args=args, intrp_dtype=np.float64, **p)
    assert rel_err(got['NNB'], exp['NNB']) == 0.0


def test_nrst_vs_oracle(eng):
    """'nrst' neighbour selection: per-cell k nearest available stations, cells
    grouped by neighbour set, small systems; OK / SK / EDK / IDW / NNB with missing
    data, a low-value step (-> neighbour mean) and a nugget-only variogram."""
    p = make_problem(31, 45, 7, 21, 17, cell=4000.0, miss=0.15)
    cx, cy = p['cell_xs'], p['cell_ys']
    p['data'][2, :] = np.where(np.isnan(p['data'][2, :]), np.nan, 0.02)
    vgs = [VG_C1] * 7
    vgs[4] = '0.2 Nug(0.0) + 0.8 Exp(30000)'
    vgs[5] = '0.0 Nug(0.0)'
    drft = np.vstack([200 + 0.003 * cx + 0.001 * cy + 30 * np.sin(cx / 9000.0)])
    sdrft = np.column_stack([200 + 0.003 * p['stn_xs'] + 0.001 * p['stn_ys']
                             + 30 * np.sin(p['stn_xs'] / 9000.0)])
    args = [('OK', None, 'OK'), ('SK', None, 'SK'), ('EDK', None, 'EDK'),
            ('IDW', None, 'IDW_000', 1.5), ('NNB', None, 'NNB')]
    kw = dict(interp_args=args, vgs=vgs, neb_sel_mthd='nrst', n_nebs=9, min_var_thr=0.1,
              min_var_cut=0.0, drft_arrs=drft, stns_drft=sdrft, **p)
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    got, _ = eng.interp_chunk(intrp_dtype=np.float64, **kw)
    _check(got, exp, 'nrst')
    assert eng.stats.get('nrst_systems', 0) > 10


def test_pie_vs_oracle(eng):
    """'pie' neighbour selection (sector round-robin, interp/grps.py:168-247): same
    estimators as 'nrst' behind a different per-cell station choice; stations placed
    exactly north / east / west of cell centres exercise the sector edge rules."""
    p = make_problem(33, 60, 6, 19, 23, cell=4000.0, miss=0.12)
    cx, cy = p['cell_xs'], p['cell_ys']
    p['stn_xs'][0], p['stn_ys'][0] = cx[40], cy[40] + 7000.0     # due north of cell 40
    p['stn_xs'][1], p['stn_ys'][1] = cx[41] - 9500.0, cy[41]     # due west of cell 41
    p['stn_xs'][2], p['stn_ys'][2] = cx[42] + 5000.0, cy[42]     # due east of cell 42
    vgs = [VG_C1] * 6
    vgs[3] = '0.2 Nug(0.0) + 0.8 Exp(30000)'
    args = [('OK', None, 'OK'), ('SK', None, 'SK'), ('IDW', None, 'IDW_000', 2.0),
            ('NNB', None, 'NNB')]
    for n_nebs, n_pies in ((9, 4), (12, 5), (8, 8)):   # (no two stations equidistant from a cell)
        kw = dict(interp_args=args, vgs=vgs, neb_sel_mthd='pie', n_nebs=n_nebs, n_pies=n_pies,
                  min_var_cut=0.0, **p)
        exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
        got, _ = eng.interp_chunk(intrp_dtype=np.float64, **kw)
        _check(got, exp, 'pie %d/%d' % (n_nebs, n_pies))
    assert eng.stats.get('nrst_systems', 0) > 10


def test_per_step_variograms_multivg(eng):
    """One variogram per step (config 3 style): the per-row-variogram estimator
    (distances of a cell tile shared by all rows) with OK, EDK and SK, missing
    data (downdated and direct systems), a mask and every variogram family."""
    p = make_problem(41, 50, 14, 23, 19, cell=3000.0, miss=0.1)
    rng = np.random.default_rng(5)
    fam = ['Sph', 'Exp', 'Gau', 'Lin', 'Hol', 'Pow']
    vgs = []
    for t in range(14):
        f = fam[t % len(fam)]
        rg = rng.uniform(1.5e4, 5e4) if f != 'Pow' else 0.5
        sill = rng.uniform(0.5, 1.5) if f != 'Pow' else 0.002
        vgs.append('%0.5f Nug(0.0) + %0.5f %s(%0.5f)' % (rng.uniform(0.15, 0.3), sill, f, rg))
    cx, cy = p['cell_xs'], p['cell_ys']
    mask = (cx + 0.7 * cy) < 6.0e4
    drft = (100 + 0.002 * cx + 0.001 * cy)[None, mask]
    sdrft = (100 + 0.002 * p['stn_xs'] + 0.001 * p['stn_ys'])[:, None]
    p['cell_xs'], p['cell_ys'] = cx[mask], cy[mask]
    args = [('OK', None, 'OK'), ('SK', None, 'SK'), ('EDK', None, 'EDK')]
    kw = dict(interp_args=args, vgs=vgs, cntn_idxs=mask, drft_arrs=drft, stns_drft=sdrft, **p)
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    got, _ = eng.interp_chunk(intrp_dtype=np.float64, **kw)
    if not eng.local_support:
        assert eng.stats.get('multivg_evals', 0) > 0
    assert eng.stats.get('pinv_systems', 0) > 0      # the Hol systems (cond > 1e16)
    for lab, ref in exp.items():
        per_step = [rel_err(got[lab][t], ref[t], _floor(ref)) for t in range(14)]
        print(lab, ['%.1e' % e for e in per_step])
        good = [per_step[t] for t in range(14) if fam[t % 6] in ('Sph', 'Exp', 'Lin', 'Pow')]
        assert max(good) <= KRG_TOL, (lab, per_step)
        # Gau systems: cond 5e8 .. 3e11, agreement ~cond * eps
        assert max(per_step[t] for t in range(14) if fam[t % 6] == 'Gau') <= 1e-3, per_step
        # Hol systems: cond > 1e16 with a singular value AT np.linalg.pinv's 1e-15
        # cut-off -- whether it is truncated depends on the last bits of the SVD, so
        # the reference itself is not reproducible there.  With the same operator
        # (pseudo-inverse, sum(lambda) test, NNB fallback) the bulk of the cells
        # (nearest-neighbour values) must still coincide.
        for t in range(14):
            if fam[t % 6] == 'Hol':
                m = np.isfinite(ref[t])
                same = np.abs(got[lab][t][m] - ref[t][m]) <= 1e-6 * np.maximum(1.0, np.abs(ref[t][m]))
                assert same.mean() >= 0.8, (lab, t, same.mean())


def test_local_estimator_is_used_and_matches_dense():
    """Compact variogram, sparse stations: the local estimator (stations within the
    range of each cell + the constant far field) must give the dense contraction's
    numbers; EDK with two drifts, SK, a mask, cut-offs, missing data."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(51, 150, 30, 60, 70, cell=3000.0, miss=0.2)
    cx, cy = p['cell_xs'], p['cell_ys']
    mask = ((cx - 1.0e5) / 9e4) ** 2 + ((cy - 9e4) / 8e4) ** 2 <= 1.0
    drft = np.vstack([100 + 0.002 * cx + 0.001 * cy, np.cos(cx / 4e4)])[:, mask]
    sdrft = np.column_stack([100 + 0.002 * p['stn_xs'] + 0.001 * p['stn_ys'],
                             np.cos(p['stn_xs'] / 4e4)])
    p['cell_xs'], p['cell_ys'] = cx[mask], cy[mask]
    vgs = ['0.1 Nug(0.0) + 0.9 Sph(20000)'] * 15 + ['0.2 Nug(0.0) + 0.5 Sph(15000) + 0.3 Lin(30000)'] * 15
    args = [('OK', None, 'OK'), ('SK', None, 'SK'), ('EDK', None, 'EDK')]
    kw = dict(interp_args=args, vgs=vgs, cntn_idxs=mask, drft_arrs=drft, stns_drft=sdrft,
              min_var_cut=0.0, max_var_cut=30.0, **p)
    e1 = ChunkEngine()
    got, _ = e1.interp_chunk(intrp_dtype=np.float64, **kw)
    assert e1.stats.get('local_rows', 0) > 0 and e1.stats.get('gemm_launches', 0) <= 2
    e2 = ChunkEngine()
    e2.local_support = False
    ref, _ = e2.interp_chunk(intrp_dtype=np.float64, **kw)
    assert e2.stats.get('local_rows', 0) == 0
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    for lab in ('OK', 'SK', 'EDK'):
        assert rel_err(got[lab], ref[lab], _floor(ref[lab])) <= 1e-11, lab
        assert rel_err(got[lab], exp[lab], _floor(exp[lab])) <= KRG_TOL, lab


@pytest.mark.parametrize('submit', ['native', 'python'])
@pytest.mark.parametrize('variant', ['local_f64', 'local_f32', 'dense', 'edk_two_vgs', 'flagged'])
def test_native_planned_fast_path_matches_general_path(variant, submit):
    """Chunks after the first of a job (full-system inverse cached) are planned by the
    native host planner and solved by one downdate launch per variogram that also emits
    the local estimator's base / transposed coefficients (engine._krige_fast); with one
    variogram per chunk the whole label is ONE native call whose solve phase runs on a
    second stream (spx_fast_submit, submit == 'native').  Same numbers as the general
    NumPy-planned path and the oracle; repeated availability patterns, a single-station
    step, a step without stations and a low-value step ride along."""
    from spinterps_b200.engine import ChunkEngine
    n_steps = 40
    p = make_problem(61, 130, n_steps, 45, 52, cell=4000.0, miss=0.15)
    rng = np.random.default_rng(62)
    data2 = rng.gamma(1.0, 5.0, size=p['data'].shape)
    data2[rng.random(data2.shape) < 0.15] = np.nan
    for t in (7, 8):                                               # same pattern as step 3
        data2[t] = np.where(np.isnan(data2[3]), np.nan, rng.gamma(1.0, 5.0, size=data2.shape[1]))
    data2[11, :] = np.nan
    data2[11, 17] = 2.5                                            # single-station step
    data2[12, :] = np.nan                                          # no station at all
    data2[13] = np.where(np.isnan(data2[13]), np.nan, 0.01)        # below min_var_thr
    vgs = [VG_C1] * n_steps
    args = [('OK', None, 'OK')]
    kw = dict(min_var_thr=0.1, min_var_cut=0.0)
    dt = np.float64
    local = True
    if variant == 'local_f32':
        dt = np.float32
    elif variant == 'dense':
        local = False
    elif variant == 'flagged':
        pass
    elif variant == 'edk_two_vgs':
        cx, cy = p['cell_xs'], p['cell_ys']
        kw['drft_arrs'] = np.vstack([100 + 0.002 * cx + 0.001 * cy])
        kw['stns_drft'] = np.column_stack([100 + 0.002 * p['stn_xs'] + 0.001 * p['stn_ys']])
        args = [('EDK', None, 'EDK'), ('OK', None, 'OK')]
        vgs = [VG_C1] * 20 + ['0.2 Nug(0.0) + 0.5 Sph(15000) + 0.3 Lin(30000)'] * 20
    lambda_tol = None
    if variant == 'flagged':
        # a zero screening tolerance flags every system: the call is redone by the general
        # path (exact per-cell sum(lambda) test)
        lambda_tol = 0.0
    kw.update(interp_args=args, vgs=vgs)
    base = {k: v for k, v in p.items() if k != 'data'}

    def run(native):
        e = ChunkEngine()
        e.local_support = local
        e.native_plan = native
        e.native_submit = native and submit == 'native'
        if lambda_tol is not None:
            e.lambda_tol = lambda_tol
        e.interp_chunk(p['data'], intrp_dtype=dt, **kw, **base)      # fills the caches
        out, prob = e.interp_chunk(data2, intrp_dtype=dt, **kw, **base)
        return e, out, prob

    e1, got, prob1 = run(True)
    e0, ref, prob0 = run(False)
    if submit == 'native' and variant != 'edk_two_vgs':
        assert e1.stats.get('native_submits', 0) == 1 and e1.stats.get('native_plans', 0) == 0
    else:
        assert e1.stats.get('native_plans', 0) >= 1 and e1.stats.get('native_submits', 0) == 0
    assert e0.stats.get('native_plans', 0) == 0 and e0.stats.get('native_submits', 0) == 0
    assert prob1 == prob0 == [12]
    if variant == 'flagged':
        assert e1.stats.get('fast_path_redo', 0) + e1.stats.get('downdate_redo', 0) >= 1
    else:
        assert e1.stats.get('fast_path_redo', 0) + e1.stats.get('downdate_redo', 0) == 0
    exp, _ = orc.interp_chunk(data2, intrp_dtype=np.float64, faithful=False, **kw, **base)
    for lab in got:
        assert np.array_equal(np.isnan(got[lab]), np.isnan(ref[lab])), lab
        if dt == np.float32:
            assert rel_err(got[lab], ref[lab], _floor(ref[lab])) <= 3e-7, lab
            assert rel_err(got[lab], exp[lab].astype(np.float32), _floor(exp[lab])) <= 3e-7, lab
        elif variant == 'flagged':
            assert rel_err(got[lab], ref[lab], _floor(ref[lab])) <= 1e-12, lab   # same path twice
        else:
            assert rel_err(got[lab], ref[lab], _floor(ref[lab])) <= 1e-11, lab
            assert rel_err(got[lab], exp[lab], _floor(exp[lab])) <= KRG_TOL, lab


@pytest.mark.parametrize('n_stn,vg,note', [
    (150, '0.1 Nug(0.0) + 0.9 Sph(20000)', 'few stations per tile: staged slices'),
    (400, '0.1 Nug(0.0) + 0.9 Sph(40000)', '> 32 distinct stations per tile: global gathers'),
])
def test_local_kernel_tile_staging_is_bit_identical(n_stn, vg, note):
    """The streamlined local kernel with the coefficient slices of a 256-cell tile staged
    in shared memory (spx_local_tiles_dev) performs the same FMAs in the same order as the
    variant that gathers from global memory: f32 fields must be identical bit for bit,
    including tiles whose station list overflows (fallback inside the same kernel), a
    ragged last tile and a ragged last row block."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(71, n_stn, 131, 61, 67, cell=3000.0, miss=0.1)
    kw = dict(interp_args=[('OK', None, 'OK')], vgs=[vg] * 131, min_var_cut=0.0, **p)
    outs = []
    for tiles in (True, False):
        e = ChunkEngine()
        e.local_tiles = tiles
        e.local_max_near = 64.0
        got, _ = e.interp_chunk(intrp_dtype=np.float32, **kw)
        assert e.stats.get('local_rows', 0) == 131
        outs.append(got['OK'])
    assert np.array_equal(outs[0], outs[1], equal_nan=True), note
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    assert rel_err(outs[0], exp['OK'].astype(np.float32), _floor(exp['OK'])) <= 3e-7


@pytest.mark.parametrize('ny,nx,n_steps', [(60, 68, 131), (64, 64, 128), (33, 36, 9)])
def test_local_kernel_bulk_stores_are_bit_identical(ny, nx, n_steps):
    """The streamlined local kernel writing through shared-memory staged row segments and
    bulk async (TMA) stores gives the same bits as the per-lane streaming stores: ragged
    last cell tile (n_cells % 256 != 0), ragged row groups (n_steps % 8 != 0, % 4 != 0),
    several row blocks."""
    from spinterps_b200 import _lib
    from spinterps_b200.engine import ChunkEngine
    assert (ny * nx) % 4 == 0
    p = make_problem(73, 150, n_steps, ny, nx, cell=3000.0, miss=0.1)
    kw = dict(interp_args=[('OK', None, 'OK')], vgs=[VG_C1] * n_steps, min_var_cut=0.5, **p)
    outs = []
    lib = _lib.load()
    prev = lib.spx_local_set_bulk(1)
    try:
        for bulk in (1, 0):
            lib.spx_local_set_bulk(bulk)
            e = ChunkEngine()
            got, _ = e.interp_chunk(intrp_dtype=np.float32, **kw)
            assert e.stats.get('local_rows', 0) == n_steps
            outs.append(got['OK'])
    finally:
        lib.spx_local_set_bulk(prev)
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    assert rel_err(outs[0], exp['OK'].astype(np.float32), _floor(exp['OK'])) <= 3e-7


@pytest.mark.parametrize('n_stn,n_rows,n_data,n_border', [(130, 77, 50, 1), (500, 300, 200, 1),
                                                           (64, 64, 64, 2), (33, 5, 3, 3)])
def test_ut_gemm_matches_matmul(n_stn, n_rows, n_data, n_border):
    """Ut = Bt . G on the FP64 tensor cores with Bt generated from the resident data block
    (spx_ut_gemm_dev) against the dense product of the explicitly built Bt."""
    import ctypes as C
    import torch
    from spinterps_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    T = 40
    data = rng.gamma(1.0, 5.0, size=(T, n_stn))
    data[rng.random(data.shape) < 0.2] = np.nan
    M = n_stn + n_border
    G = rng.normal(size=(M, M))
    G = G + G.T
    src = rng.integers(0, T, size=n_rows).astype(np.int32)
    bt = np.zeros((n_rows, M))
    fin = np.isfinite(data[src])
    bt[:n_data, :n_stn] = np.where(fin[:n_data], data[src[:n_data]], 0.0)
    bt[n_data:, :n_stn] = fin[n_data:].astype(float)
    d_data = torch.from_numpy(data).cuda()
    d_src = torch.from_numpy(src).cuda()
    d_G = torch.from_numpy(G).cuda()
    ut = torch.full((n_rows, M), float('nan'), dtype=torch.float64, device='cuda')
    _lib.check(lib.spx_ut_gemm_dev(
        C.c_void_p(d_data.data_ptr()), n_stn, n_stn, C.c_void_p(d_src.data_ptr()), n_rows, n_data,
        n_border, C.c_void_p(d_G.data_ptr()), C.c_void_p(ut.data_ptr()),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'ut_gemm')
    torch.cuda.synchronize()
    exp = bt @ G
    scale = np.abs(bt) @ np.abs(G) + 1e-300
    assert np.max(np.abs(ut.cpu().numpy() - exp) / scale) <= 1e-14


def test_native_submit_pipeline_many_chunks():
    """Several chunks in flight through the native submit (ring of slots, solve stream
    overlapping the previous estimate): every chunk's field equals the one computed alone by
    the Python-planned path; chunk sizes vary (ragged last chunk)."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(81, 120, 64, 40, 50, cell=4000.0, miss=0.2)
    base = {k: v for k, v in p.items() if k != 'data'}
    kw = dict(interp_args=[('OK', None, 'OK')], intrp_dtype=np.float32, **base)
    rng = np.random.default_rng(82)
    chunks = []
    for n in (64, 64, 30, 64, 17, 64, 64, 5, 64):
        d = rng.gamma(1.0, 5.0, size=(n, 120))
        d[rng.random(d.shape) < 0.2] = np.nan
        chunks.append(d)
    e1 = ChunkEngine()
    e1.interp_chunk(chunks[0], vgs=[VG_C1] * 64, **kw)           # caches
    pend = [e1.submit_chunk(d, vgs=[VG_C1] * d.shape[0], **kw) for d in chunks]
    outs = [pd.result()[0]['OK'] for pd in pend]
    assert e1.stats.get('native_submits', 0) == 1
    e0 = ChunkEngine()
    e0.native_submit = False
    e0.interp_chunk(chunks[0], vgs=[VG_C1] * 64, **kw)
    for d, got in zip(chunks, outs):
        ref, _ = e0.interp_chunk(d, vgs=[VG_C1] * d.shape[0], **kw)
        assert got.shape == ref['OK'].shape
        assert rel_err(got, ref['OK'], _floor(ref['OK'])) <= 3e-7
    e1.close()


def _assert_same_neighbours(eng, mthd, n_nebs, n_pies, sx, sy, cx, cy, avail=None):
    got_idx, got_grps = eng.neighbor_indices(sx, sy, cx, cy, mthd, n_nebs, n_pies=n_pies,
                                             avail=avail)
    if avail is None:
        exp_idx, exp_grps = orc.get_neb_idxs_and_grps(mthd, n_nebs, cx, cy, sx, sy, n_pies=n_pies)
    else:
        keep = np.where(avail)[0]
        loc_idx, exp_grps = orc.get_neb_idxs_and_grps(mthd, n_nebs, cx, cy, sx[keep], sy[keep],
                                                      n_pies=n_pies)
        exp_idx = keep[loc_idx]
    assert got_idx.dtype == np.int64 and got_idx.shape == exp_idx.shape
    assert np.array_equal(got_idx, exp_idx)                       # bit-exact index rows
    assert len(got_grps) == len(exp_grps)
    for a, b in zip(got_grps, exp_grps):                          # same groups, same order
        assert np.array_equal(a, b)


@pytest.mark.parametrize('name', ['e_nrst', 'h_pie'])
def test_neighbour_and_cell_group_indices_bit_exact_on_golden_inputs(name):
    """north_star: station grouping, neighbour and cell-selection indices bit-exact.  The
    index arrays of spx_nrst_topk_dev / spx_pie_select_dev and the neighbour-hash cell
    groups are compared DIRECTLY (not through the fields) with the oracle's
    get_neb_idxs_and_grps (interp/grps.py:103-288) on the inputs of the reference-generated
    golden cases, for all stations and for one availability subset."""
    from spinterps_b200.engine import ChunkEngine
    case, _ = load_case(name)
    e = ChunkEngine()
    mthd, k, npies = case['neb_sel_mthd'], case['n_nebs'], case.get('n_pies')
    sx, sy, cx, cy = case['stn_xs'], case['stn_ys'], case['cell_xs'], case['cell_ys']
    _assert_same_neighbours(e, mthd, k, npies, sx, sy, cx, cy)
    avail = np.isfinite(case['data'][0])
    if avail.sum() > k and not avail.all():
        _assert_same_neighbours(e, mthd, k, npies, sx, sy, cx, cy, avail=avail)


@pytest.mark.parametrize('mthd,n_stn,k,n_pies', [('nrst', 2000, 50, None), ('nrst', 1000, 64, None),
                                                  ('pie', 300, 12, 4)])
def test_neighbour_indices_bit_exact_large_station_sets(mthd, n_stn, k, n_pies):
    """The reference's canonical setting (test/test_interp.py:101-103: nrst, 50
    neighbours) at N = 2000 stations."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(93, n_stn, 2, 40, 50, cell=2500.0)
    rng = np.random.default_rng(94)
    avail = rng.random(n_stn) > 0.2
    e = ChunkEngine()
    _assert_same_neighbours(e, mthd, k, n_pies, p['stn_xs'], p['stn_ys'], p['cell_xs'],
                            p['cell_ys'])
    _assert_same_neighbours(e, mthd, k, n_pies, p['stn_xs'], p['stn_ys'], p['cell_xs'],
                            p['cell_ys'], avail=avail)


@pytest.mark.parametrize('n_stn,k,lattice', [(77, 10, True), (500, 100, False), (1000, 50, True),
                                              (2500, 160, False), (33, 40, False)])
def test_topk_warp_kernel_equals_the_insertion_kernel(n_stn, k, lattice):
    """spx_nrst_topk_dev has a warp-per-cell kernel (bisection on the distance keys) and a
    thread-per-cell kernel (sorted insertion): rows AND hashes identical, with equidistant
    stations (lattice coordinates: ties go to the lower index), masks, k > 64, k > number
    of available stations (padded with -1) and a station count that is not a multiple of 32."""
    import torch
    from spinterps_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n_stn + k)
    if lattice:
        sx = rng.integers(0, 12, n_stn).astype(np.float64) * 1000.0
        sy = rng.integers(0, 12, n_stn).astype(np.float64) * 1000.0
        cx = rng.integers(0, 24, 3000).astype(np.float64) * 500.0
        cy = rng.integers(0, 24, 3000).astype(np.float64) * 500.0
    else:
        sx, sy = rng.uniform(0, 1e5, n_stn), rng.uniform(0, 1e5, n_stn)
        cx, cy = rng.uniform(0, 1e5, 3001), rng.uniform(0, 1e5, 3001)
    dev = torch.device('cuda:0')
    d_sx, d_sy = torch.from_numpy(sx).to(dev), torch.from_numpy(sy).to(dev)
    d_cx, d_cy = torch.from_numpy(cx).to(dev), torch.from_numpy(cy).to(dev)
    masks = [None, torch.from_numpy((rng.random(n_stn) > 0.3).astype(np.uint8)).to(dev)]
    prev = lib.spx_nrst_set_topk_warp(1)
    try:
        for d_mask in masks:
            res = []
            for warp in (1, 0):
                lib.spx_nrst_set_topk_warp(warp)
                nb = torch.full((cx.size, k), -7, dtype=torch.int32, device=dev)
                hsh = torch.full((cx.size,), -7, dtype=torch.int64, device=dev)
                _lib.check(lib.spx_nrst_topk_dev(
                    d_sx.data_ptr(), d_sy.data_ptr(), n_stn,
                    None if d_mask is None else d_mask.data_ptr(), d_cx.data_ptr(),
                    d_cy.data_ptr(), cx.size, k, nb.data_ptr(), hsh.data_ptr(), None), 'topk')
                torch.cuda.synchronize()
                res.append((nb.cpu().numpy(), hsh.cpu().numpy()))
            assert np.array_equal(res[0][0], res[1][0])
            assert np.array_equal(res[0][1], res[1][1])
            # against NumPy: the k smallest (distance, index) pairs, indices ascending
            ok = np.ones(n_stn, bool) if d_mask is None else d_mask.cpu().numpy().astype(bool)
            for c in rng.integers(0, cx.size, 25):
                dx, dy = cx[c] - sx, cy[c] - sy
                d = np.sqrt(dx * dx + dy * dy)
                cand = np.flatnonzero(ok)
                sel = np.sort(cand[np.lexsort((cand, d[cand]))][:k])
                exp = np.concatenate([sel, np.full(k - sel.size, -1)])
                assert np.array_equal(res[0][0][c], exp), c
    finally:
        lib.spx_nrst_set_topk_warp(prev)


def test_cached_geometry_survives_the_upload_arena_ring():
    """Seven chunks through ONE engine (the upload arenas are a ring of four): every label of
    every chunk equals the result of a fresh engine.  (Device copies that outlive a chunk --
    cell coordinates, bin tables -- must not live in a per-chunk arena.)"""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(95, 40, 35, 31, 39, cell=1500.0, miss=0.15)
    base = {k: v for k, v in p.items() if k != 'data'}
    args = [('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0), ('NNB', None, 'NNB')]
    eng = ChunkEngine()
    eng_dense = ChunkEngine()
    eng_dense.local_support = False
    for i in range(7):
        d = p['data'][5 * i:5 * i + 5]
        kw = dict(interp_args=args, vgs=[VG_C1] * 5, intrp_dtype=np.float64, **base)
        ref, _ = ChunkEngine().interp_chunk(d, **kw)
        for e in (eng, eng_dense):
            got, _ = e.interp_chunk(d, **kw)
            for lab in ref:
                assert rel_err(got[lab], ref[lab], _floor(ref[lab])) <= 1e-11, (i, lab)


def test_geometry_cache_follows_in_place_coordinate_edits():
    """The cross-chunk caches are keyed by content: shifting the coordinate arrays IN PLACE
    (same buffers, same addresses) must give the fields of the shifted grid."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(96, 60, 6, 20, 25, cell=2000.0, miss=0.1)
    kw = dict(interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)], vgs=[VG_C1] * 6,
              intrp_dtype=np.float64)
    eng = ChunkEngine()
    a, _ = eng.interp_chunk(**kw, **p)
    p['cell_xs'] += 700.0                     # in place
    p['cell_ys'] -= 300.0
    b, _ = eng.interp_chunk(**kw, **p)
    ref, _ = ChunkEngine().interp_chunk(**kw, **p)
    for lab in ref:
        assert rel_err(b[lab], ref[lab], _floor(ref[lab])) <= 1e-11, lab
        assert rel_err(a[lab], ref[lab], _floor(ref[lab])) > 1e-6, lab   # the grids do differ


def test_parity_at_1000_stations_sk_ok_mask():
    """Config-5 shape at a grid the oracle finishes in seconds: 1,000 stations, SK + OK,
    elliptic cell mask, missing data (several availability groups of ~900 stations)."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(105, 1000, 6, 30, 40, cell=100000.0 / 40, miss=0.1)
    cx, cy = p['cell_xs'], p['cell_ys']
    mask = ((cx - 5.0e4) / 4.0e4) ** 2 + ((cy - 3.7e4) / 3.0e4) ** 2 <= 1.0
    assert 0.3 < mask.mean() < 0.8
    p['cell_xs'], p['cell_ys'] = cx[mask], cy[mask]
    kw = dict(interp_args=[('OK', None, 'OK'), ('SK', None, 'SK')], vgs=[VG_C1] * 6,
              cntn_idxs=mask, intrp_dtype=np.float64, **p)
    exp, _ = orc.interp_chunk(faithful=False, **kw)
    for local in (True, False):
        e = ChunkEngine()
        e.local_support = local
        got, _ = e.interp_chunk(**kw)
        _check(got, exp, 'n1000_sk_ok_mask_%s' % ('local' if local else 'dense'))
        assert np.array_equal(np.isnan(got['OK']), np.isnan(exp['OK']))


def test_parity_at_2000_stations_idw_four_exponents():
    """Config-4 shape: 2,000 stations, IDW exponents 1, 2, 3, 5 (kpad 2000: the contraction
    keeps 8-16 cells per tile resident), missing data."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(104, 2000, 5, 20, 25, cell=4000.0, miss=0.1)
    args = [('IDW', None, 'IDW_%03d' % i, float(e)) for i, e in enumerate((1, 2, 3, 5))]
    kw = dict(interp_args=args, intrp_dtype=np.float64, **p)
    exp, _ = orc.interp_chunk(faithful=False, **kw)
    got, _ = ChunkEngine().interp_chunk(**kw)
    _check(got, exp, 'n2000_idw_x4')


@pytest.mark.parametrize('min_vg_val', [0.15, 0.6])
def test_min_vg_val_through_the_local_estimator(min_vg_val):
    """min_vg_val > 0 (cyth/interpmthds.pyx:203-216: variogram values <= min_vg_val become 0,
    in A and in the right-hand sides) through the local estimator, the dense contraction and
    the oracle."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(106, 90, 8, 33, 37, cell=2500.0, miss=0.15)
    kw = dict(interp_args=[('OK', None, 'OK')], vgs=[VG_C1] * 8, min_vg_val=min_vg_val,
              intrp_dtype=np.float64, **p)
    exp, _ = orc.interp_chunk(faithful=False, **kw)
    e1 = ChunkEngine()
    got, _ = e1.interp_chunk(**kw)
    assert e1.stats.get('local_rows', 0) > 0
    e2 = ChunkEngine()
    e2.local_support = False
    ref, _ = e2.interp_chunk(**kw)
    _check(got, exp, 'min_vg_val_%s_local' % min_vg_val)
    _check(ref, exp, 'min_vg_val_%s_dense' % min_vg_val)


@pytest.mark.parametrize('mthd,n_nebs,n_pies', [('nrst', 9, None), ('pie', 12, 4)])
def test_est_vars_ok_with_nrst_and_pie(mthd, n_nebs, n_pies):
    """EST_VARS_OK with per-cell neighbour selection (the reference computes it for any
    neighbour method, interp/steps.py:428-434): sum(lambda * rhs) + lambda[n] from the
    inverse of each cell group's system; 0 where NNB / the neighbour mean was written."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(35, 50, 6, 18, 22, cell=4000.0, miss=0.15)
    p['data'][2, :] = np.where(np.isnan(p['data'][2, :]), np.nan, 0.02)
    if mthd == 'nrst':      # ('pie' with fewer stations than neighbours raises in the reference)
        p['data'][4, 1:] = np.nan                               # single-station step
    args = [('OK', None, 'OK'), ('EST_VARS_OK', None, 'EST_VARS_OK')]
    kw = dict(interp_args=args, vgs=[VG_C1] * 6, neb_sel_mthd=mthd, n_nebs=n_nebs, n_pies=n_pies,
              min_var_thr=0.1, est_var_flag=True, **p)
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    got, _ = ChunkEngine().interp_chunk(intrp_dtype=np.float64, **kw)
    _check(got, exp, 'est_vars_' + mthd)
    assert np.nanmax(exp['EST_VARS_OK']) > 0.1 and (exp['EST_VARS_OK'] == 0).any()


@pytest.mark.parametrize('n_steps', [5, 150])
def test_nrst_solve_thread_and_warp_substitutions_agree(n_steps):
    """k_nrst_solve substitutes either with one thread or with one warp per right-hand side
    (same operations in the same order): identical OK / EDK / EST_VARS_OK fields; 150 steps
    need two batches of right-hand sides."""
    from spinterps_b200 import _lib
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(37, 90, n_steps, 16, 21, cell=4000.0, miss=0.1)
    rng = np.random.default_rng(38)
    drft = rng.normal(500.0, 100.0, size=(1, p['cell_xs'].size))
    sdrft = rng.normal(500.0, 100.0, size=(90, 1))
    args = [('OK', None, 'OK'), ('EDK', None, 'EDK'), ('EST_VARS_OK', None, 'EST_VARS_OK')]
    kw = dict(interp_args=args, vgs=[VG_C1] * n_steps, neb_sel_mthd='nrst', n_nebs=12,
              est_var_flag=True, drft_arrs=drft, stns_drft=sdrft, intrp_dtype=np.float64, **p)
    lib = _lib.load()
    prev = lib.spx_nrst_set_thread_rhs(1)
    try:
        outs = []
        for mode in (1, 0):
            lib.spx_nrst_set_thread_rhs(mode)
            got, _ = ChunkEngine().interp_chunk(**kw)
            outs.append(got)
    finally:
        lib.spx_nrst_set_thread_rhs(prev)
    for lab in ('OK', 'EDK', 'EST_VARS_OK'):
        assert np.isfinite(outs[0][lab]).mean() > 0.9
        assert np.array_equal(outs[0][lab], outs[1][lab], equal_nan=True), lab
    if n_steps == 5:
        exp, _ = orc.interp_chunk(faithful=False, **kw)
        _check(outs[0], {lab: exp[lab] for lab in ('OK', 'EST_VARS_OK')}, 'nrst_thread_rhs')


def test_nrst_with_100_neighbours():
    """More than 64 neighbours per cell (the reference has no cap, interp/grps.py:147-166):
    the 160-neighbour kernel variants, index rows bit-exact, OK / IDW fields vs the oracle."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(36, 260, 4, 12, 15, cell=5000.0, miss=0.1)
    e = ChunkEngine()
    _assert_same_neighbours(e, 'nrst', 100, None, p['stn_xs'], p['stn_ys'], p['cell_xs'],
                            p['cell_ys'])
    _assert_same_neighbours(e, 'pie', 100, 70, p['stn_xs'], p['stn_ys'], p['cell_xs'],
                            p['cell_ys'])
    kw = dict(interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)], vgs=[VG_C1] * 4,
              neb_sel_mthd='nrst', n_nebs=100, **p)
    exp, _ = orc.interp_chunk(intrp_dtype=np.float64, faithful=False, **kw)
    got, _ = e.interp_chunk(intrp_dtype=np.float64, **kw)
    _check(got, exp, 'nrst_100')


def test_per_step_compact_variograms_share_transient_tables():
    """Config-3 style: one compact variogram PER STEP (more than the table cache holds): the
    neighbour tables of every variogram are built into shared buffers sized once from the
    largest range; EDK with a drift and OK against the oracle, and against the dense path."""
    from spinterps_b200.engine import ChunkEngine
    T = 14
    p = make_problem(107, 80, T, 30, 34, cell=3000.0, miss=0.1)
    rng = np.random.default_rng(108)
    vgs = ['%0.5f Nug(0.0) + %0.5f Sph(%0.5f)' % (rng.uniform(0, 0.2), rng.uniform(0.5, 1.5),
                                                    rng.uniform(8e3, 2.5e4)) for _ in range(T)]
    cx, cy = p['cell_xs'], p['cell_ys']
    drft = np.vstack([300 + 0.002 * cx + 0.001 * cy])
    sdrft = np.column_stack([300 + 0.002 * p['stn_xs'] + 0.001 * p['stn_ys']])
    kw = dict(interp_args=[('EDK', None, 'EDK'), ('OK', None, 'OK')], vgs=vgs, drft_arrs=drft,
              stns_drft=sdrft, intrp_dtype=np.float64, **p)
    exp, _ = orc.interp_chunk(faithful=False, **kw)
    e1 = ChunkEngine()
    got, _ = e1.interp_chunk(**kw)
    assert e1.stats.get('local_rows', 0) == 2 * T        # every step through the local estimator
    _check(got, exp, 'per_step_compact_vgs')
    e2 = ChunkEngine()
    e2.local_support = False
    ref, _ = e2.interp_chunk(**kw)
    for lab in ref:
        assert rel_err(got[lab], ref[lab], _floor(ref[lab])) <= 1e-11, lab


def test_idw_station_on_a_cell_centre():
    """A station exactly on a cell centre has weight inf (quirk Q7): NaN at that cell for the
    steps where the station is available, and -- because the reference excludes missing
    stations before it forms the weights (interp/steps.py:293-313) -- a finite value where it
    is missing (0 * inf must not leak in)."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(109, 30, 8, 12, 14, cell=3000.0, miss=0.0)
    p['stn_xs'][5], p['stn_ys'][5] = p['cell_xs'][40], p['cell_ys'][40]
    p['data'][2:5, 5] = np.nan                       # missing for three steps
    p['data'][6, [1, 7]] = np.nan                    # another availability group
    args = [('IDW', None, 'IDW_000', 2.0), ('IDW', None, 'IDW_001', 3.0)]
    kw = dict(interp_args=args, intrp_dtype=np.float64, **p)
    with np.errstate(all='ignore'):
        exp, _ = orc.interp_chunk(faithful=False, **kw)
    got, _ = ChunkEngine().interp_chunk(**kw)
    for lab in exp:
        assert np.isnan(exp[lab][[0, 1, 5, 6, 7], 40]).all() and np.isfinite(exp[lab][2:5, 40]).all()
        assert np.isfinite(exp[lab][:, np.arange(168) != 40]).all()
    _check(got, exp, 'idw_station_on_cell_centre')


@pytest.mark.parametrize('vg', ['0.1 Nug(0.0) + 0.9 Sph(9000)',
                                '0.2 Nug(0.0) + 0.5 Sph(6000) + 0.3 Lin(11000)'])
@pytest.mark.parametrize('dt', [np.float64, np.float32])
def test_sparse_covariance_solve_matches_downdate_and_oracle(vg, dt):
    """Ordinary kriging with a compact variogram whose stations form small clusters: the
    native submit solves the block-diagonal covariance form F 11' - C in O(n_stn) per step
    (spx_krige_sparse_ok_dev) instead of downdating dense systems.  Same numbers as the
    downdate path and the oracle; single-station / empty / low-value steps ride along."""
    from spinterps_b200.engine import ChunkEngine
    n_steps = 48
    p = make_problem(71, 220, n_steps, 64, 70, cell=4000.0, miss=0.2)
    rng = np.random.default_rng(72)
    data2 = rng.gamma(1.0, 5.0, size=p['data'].shape)
    data2[rng.random(data2.shape) < 0.2] = np.nan
    data2[5, :] = np.nan
    data2[5, 40] = 3.5                                             # single-station step
    data2[6, :] = np.nan                                           # no station at all
    data2[9] = np.where(np.isnan(data2[9]), np.nan, 0.01)          # below min_var_thr
    kw = dict(min_var_thr=0.1, min_var_cut=0.0, interp_args=[('OK', None, 'OK')],
              vgs=[vg] * n_steps)
    base = {k: v for k, v in p.items() if k != 'data'}

    first = {}

    def run(sparse):
        e = ChunkEngine()
        e.sparse_solve = sparse
        out0, _ = e.interp_chunk(p['data'], intrp_dtype=dt, **kw, **base)   # fills the caches
        first[sparse] = (dict(e.stats), out0['OK'])
        out, prob = e.interp_chunk(data2, intrp_dtype=dt, **kw, **base)
        return e, out, prob

    e1, got, prob1 = run(True)
    e0, ref, prob0 = run(False)
    # the sparse form needs no inverse of the full station system: already the FIRST chunk of
    # a job goes through the native submit; without it the general path runs first
    assert first[True][0].get('native_submits', 0) == 1 and first[True][0].get('sparse_cov_jobs', 0) == 1
    assert first[False][0].get('native_submits', 0) == 0
    assert rel_err(first[True][1], first[False][1], _floor(first[False][1])) <= (3e-7 if dt == np.float32 else 1e-10)
    assert e1.stats.get('native_submits', 0) == 1
    assert any(j['cfg'].sparse.n_comp > 0 for j in e1._fast_jobs.values())
    assert e0.stats.get('native_submits', 0) == 1
    assert not any(j['cfg'].sparse.n_comp > 0 for j in e0._fast_jobs.values())
    assert e1.stats.get('fast_path_redo', 0) + e1.stats.get('downdate_redo', 0) == 0
    assert prob1 == prob0 == [6]
    exp, _ = orc.interp_chunk(data2, intrp_dtype=np.float64, faithful=False, **kw, **base)
    g, r, x = got['OK'], ref['OK'], exp['OK']
    assert np.array_equal(np.isnan(g), np.isnan(r))
    if dt == np.float32:
        assert rel_err(g, r, _floor(r)) <= 3e-7
        assert rel_err(g, x.astype(np.float32), _floor(x)) <= 3e-7
    else:
        assert rel_err(g, r, _floor(r)) <= 1e-10
        assert rel_err(g, x, _floor(x)) <= KRG_TOL


def test_sparse_covariance_solve_needs_small_clusters():
    """Stations that chain up within the range (one big component) keep the downdate."""
    from spinterps_b200.engine import ChunkEngine
    p = make_problem(73, 150, 30, 40, 44, cell=4000.0, miss=0.15)
    kw = dict(interp_args=[('OK', None, 'OK')], vgs=[VG_C1] * 30)
    base = {k: v for k, v in p.items() if k != 'data'}
    e = ChunkEngine()
    e.interp_chunk(p['data'], intrp_dtype=np.float64, **kw, **base)
    got, _ = e.interp_chunk(p['data'][::-1].copy(), intrp_dtype=np.float64, **kw, **base)
    assert e.stats.get('native_submits', 0) == 1 and e.stats.get('sparse_cov_jobs', 0) == 0
    exp, _ = orc.interp_chunk(p['data'][::-1].copy(), intrp_dtype=np.float64, faithful=False,
                              **kw, **base)
    assert rel_err(got['OK'], exp['OK'], _floor(exp['OK'])) <= KRG_TOL


@pytest.mark.parametrize('kind', ['idw', 'ok'])
def test_gemm_k_split_matches_single_pass(kind):
    """Large K (>~ 1000 stations): the contraction runs in several passes over K whose partial
    sums travel through HBM (spx_gemm_set_ksplit); same numbers as one pass over the full K
    (different summation order: 1e-13), incl. a station on a cell centre that is found in a
    pass other than the last one, missing data and a ragged number of cells."""
    from spinterps_b200 import _lib
    from spinterps_b200.engine import ChunkEngine
    lib = _lib.load()
    p = make_problem(131, 1300, 6, 33, 37, cell=3000.0, miss=0.1)
    p['stn_xs'][7], p['stn_ys'][7] = p['cell_xs'][400], p['cell_ys'][400]     # first K pass
    p['data'][:, 7] = 4.5
    p['data'][2, 7] = np.nan
    if kind == 'idw':
        args = [('IDW', None, 'IDW_000', 2.0), ('IDW', None, 'IDW_001', 3.0)]
        kw = dict(interp_args=args, intrp_dtype=np.float64, **p)
    else:
        kw = dict(interp_args=[('OK', None, 'OK')], vgs=['0.1 Nug(0.0) + 0.9 Exp(60000)'] * 6,
                  intrp_dtype=np.float64, **p)
    outs = {}
    prev = lib.spx_gemm_set_ksplit(4)
    try:
        for passes in (4, 0):
            lib.spx_gemm_set_ksplit(passes)
            e = ChunkEngine()
            e.local_support = False
            outs[passes], _ = e.interp_chunk(**kw)
    finally:
        lib.spx_gemm_set_ksplit(prev)
    for lab in outs[0]:
        a, b = outs[4][lab], outs[0][lab]
        assert np.array_equal(np.isnan(a), np.isnan(b)), lab
        assert rel_err(a, b, _floor(b)) <= 1e-12, lab
    if kind == 'idw':
        assert np.isnan(outs[4]['IDW_000'][[0, 1, 3, 4, 5], 400]).all()
        assert np.isfinite(outs[4]['IDW_000'][2, 400])
